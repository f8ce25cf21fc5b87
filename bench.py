"""Benchmark of the hot path: CoR2 (or ODA) fwd + KLD loss + bwd, samples/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--model CoR2|ODA] [--batch 256] [--regions 36] [--precision fp32|tf32x3|tf32|bf16]

One JSON line on stdout (rank 0).  Workload at N=1: BASELINE.json configs[1], "CoR2 forward+backward
on 1 B200, batch 256 x 36 regions x 2048-d, fp32 parity mode" (the stock config/CoR2.py chain:
att1 -> compound objects -> att2), train mode (dropout on), batch 256 PER GPU (weak scaling).
  value     device-resident inputs, timed region = K steps bracketed by CUDA events
  e2e       same step through model(sample) with PINNED HOST inputs: H2D of v/q/a and a D2H read of the
            loss inside the timed region, every step
  roofline  the op with the largest share of the step (per-op CUDA events, vqa_profile_*), against
            MEASURED_PEAKS.json
  cpu_baseline  the oracle (restated reference, torch CPU) on this box's host cores, bounded sample
`--impl reference` times that CPU oracle alone (the reference is pure Python/ATen; /root/reference does
not exist on the GPU box, so the arm runs the committed restatement, kind "port").
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

H, F, A, Q, D, G = 310, 510, 620, 2400, 2048, 4
NUM_ANS = {"CoR2": 2000, "ODA": 3000}


# ----------------------------------------------------------------------------------------- helpers
def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="CoR2", choices=["CoR2", "ODA"])
    ap.add_argument("--batch", type=int, default=256, help="per GPU")
    ap.add_argument("--regions", type=int, default=36)
    ap.add_argument("--precision", default="bf16x3",
                    help="fp32-parity modes: bf16x3 (default: bf16 hi+lo operand planes, 3 MMAs), tf32x3, fp32 (CUDA cores); "
                         "reduced precision: bf16, tf32")
    ap.add_argument("--cpu-batch", type=int, default=0,
                    help="batch of the CPU oracle: default = --batch for `--impl reference` (shrunk only if host RAM "
                         "cannot hold the reference's materialised tensors), 64 for the cpu_baseline leg of the GPU arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eval-mode", action="store_true", help="dropout off (default: train mode)")
    ap.add_argument("--no-overlap", action="store_true",
                    help="N>1: run the gradient all-reduce after the graph replay instead of capturing it, bucket by "
                         "bucket, inside the backward")
    ap.add_argument("--transport", default="auto", choices=["auto", "peer", "nccl"],
                    help="N>1 gradient all-reduce: libvqacore's NVLink peer-memory kernel (peer), NCCL, or auto = peer "
                         "when symmetric memory can be set up")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm": d["hbm_gbs"], "tensor_burst": d["bf16_tflops"], "tensor": d.get("bf16_tflops_sustained",
                d["bf16_tflops"]), "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake_slowdown": 0x80, "sw_power_cap": 0x4}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def op_work(model, B, N, C):
    """Algorithmic work per launch of each plan op: ('tensor', flops) or ('hbm', bytes)."""
    M = B * N
    gemm = lambda m, k, n: 2.0 * m * k * n
    w = {
        "compress_v.fwd": ("tensor", gemm(M, D, H)), "compress_v2.fwd": ("tensor", gemm(M, D, H)),
        "compress_v.bwd": ("tensor", gemm(M, D, H)), "compress_v2.bwd": ("tensor", 2 * gemm(M, D, H)),
        "fusion_vq1.fwd": ("tensor", 2 * gemm(M, H, F)), "fusion_vq2.fwd": ("tensor", 2 * gemm(M, H, F)),
        "fusion_vq1.bwd": ("tensor", 4 * gemm(M, H, F)), "fusion_vq2.bwd": ("tensor", 4 * gemm(M, H, F)),
        "att1.pool.fwd": ("hbm", 4.0 * (M * D + M * F + B * G * D + M * G)),
        "att2.pool.fwd": ("hbm", 4.0 * (M * D + M * F + B * G * D + M * G)),
        "att1.pool.bwd": ("hbm", 4.0 * (M * D + 2 * M * F)), "att2.pool.bwd": ("hbm", 4.0 * (2 * M * D + 2 * M * F)),
        "compound.fwd": ("hbm", 4.0 * 2 * M * D), "compound.bwd": ("hbm", 4.0 * 2 * M * D),
        "oda_pair_attn.fwd": ("hbm", 4.0 * (M * D + M * H)), "oda_pair_attn.bwd": ("hbm", 4.0 * (M * D + 2 * M * H)),
        "q_proj4.fwd": ("tensor", 4 * gemm(B, Q, H)), "q_proj4.bwd": ("tensor", 4 * gemm(B, Q, H)),
        "q_proj2.fwd": ("tensor", 2 * gemm(B, Q, H)), "q_proj2.bwd": ("tensor", 2 * gemm(B, Q, H)),
        "gates.fwd": ("tensor", 2 * gemm(B, H, D)), "gates.bwd": ("tensor", 4 * gemm(B, H, D)),
        "classif.fwd": ("tensor", gemm(B, F, C)), "classif.bwd": ("tensor", 2 * gemm(B, F, C)),
    }
    R, K1 = (2, 2 * A) if model == "CoR2" else (5, A)
    w["fusion_final.fwd"] = ("tensor", R * (gemm(B, K1, F) + gemm(B, H, F)))
    w["fusion_final.bwd"] = ("tensor", 2 * R * (gemm(B, K1, F) + gemm(B, H, F)))
    for a in ("att1", "att2", "att"):
        w[a + ".glimpse.fwd"] = ("tensor", 4 * gemm(B, D, A // G))
        w[a + ".glimpse.bwd"] = ("tensor", 8 * gemm(B, D, A // G))
    return w


MATH_NOTE = {
    "tf32x3": "fp32-parity math: 3 TF32 tensor-core passes per product (x = hi + lo), so at most 1/6 of the dense bf16 "
              "peak is reachable",
    "bf16x3": "fp32-parity math: 3 bf16 tensor-core passes per product (x = hi + lo planes), so at most 1/3 of the "
              "dense bf16 peak is reachable",
    "tf32": "single-pass TF32 tensor-core math: at most 1/2 of the dense bf16 peak is reachable",
    "bf16": "bf16 tensor-core math, fp32 accumulate",
    "fp32": "fp32 CUDA-core math",
}


def kernel_traffic(label):
    """DRAM bytes per launch of a kernel from the committed ncu capture (profiles/r*_kernel_traffic.json), or None."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    for fn in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if fn.endswith("_kernel_traffic.json"):
            try:
                v = json.load(open(os.path.join(pdir, fn))).get(label)
            except Exception:
                v = None
            if v is not None:
                best = v               # later rounds override earlier ones
    return best


def kernel_roofline(per_kernel, per_op, total, pk, args, B, N, C):
    """Roofline of the dominant KERNEL of the step: the launch with the largest average duration over EVERY kernel the
    C ABI records while profiling ('k:<what> M.. N.. K.. g.. s..' for the tensor-core GEMMs, 'k:<name> hbm=<bytes>' for
    the bandwidth-bound kernels, 'k:<name> flop=<fp32 flops>' for the CUDA-core pairwise kernels), timed live with
    CUDA events on the launching stream.  GEMM: achieved = algorithmic flops (2*M*N*K per group, ONE pass — the extra
    passes of an fp32-parity mode are an implementation cost, not work) / duration against the measured dense bf16
    rate.  HBM kernels: algorithmic bytes / duration against the measured copy bandwidth.  traffic = DRAM bytes of the
    same launch from the committed ncu capture (profiles/), or null."""
    share = lambda ms: ms / total if total else None
    if per_kernel:
        top = max(per_kernel, key=per_kernel.get)
        label, ms = top[2:], per_kernel[top]
        dur = ms * 1e-3
        traffic = kernel_traffic(label)
        if " hbm=" in label or " flop=" in label:
            name, work = label.rsplit(" ", 1)
            kind, amount = work.split("=")
            amount = float(amount)
            if kind == "hbm":
                ach = amount / dur / 1e9
                return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
                        "frac": ach / pk["hbm"], "traffic": traffic, "ms": ms, "share_of_step": share(ms),
                        "algorithmic_bytes": amount, "note": "peak = measured copy bandwidth, %s" % pk["src"]}
            # CUDA-core kernel (ODA train-mode pairwise terms): neither HBM- nor tensor-bound; the schema's nearest
            # roof is the tensor one, so the fraction is quoted against the fp32 FMA rate and says so
            ach = amount / dur / 1e12
            fp32_peak = 2.0 * 128 * 148 * 1.965e9 / 1e12       # 128 FMA lanes/SM x 148 SMs x max clock
            return {"kernel": name, "bound": "tensor", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s",
                    "frac": ach / fp32_peak, "traffic": traffic, "ms": ms, "share_of_step": share(ms),
                    "algorithmic_flops": amount,
                    "note": "CUDA-core kernel (per-element Philox mask in registers, output width 4: no tensor-core "
                            "form); peak = fp32 FMA issue rate 2*128*148*1.965 GHz, not the tensor peak"}
        dims = {t[0]: int(t[1:]) for t in label.split()[1:]}
        flops = 2.0 * dims["M"] * dims["N"] * dims["K"] * dims["g"]
        ach = flops / dur / 1e12
        mode = label.split()[0].rsplit("@", 1)[1] if "@" in label.split()[0] else args.precision
        return {"kernel": label, "bound": "tensor", "achieved": ach, "peak": pk["tensor"], "unit": "TFLOP/s",
                "frac": ach / pk["tensor"], "traffic": traffic, "ms": ms, "share_of_step": share(ms),
                "algorithmic_flops": flops,
                "note": "%s; peak = measured dense bf16 sustained, %s" % (MATH_NOTE.get(mode, mode), pk["src"])}
    work = op_work(args.model, B, N, C)
    top = max(per_op, key=per_op.get)
    kind, amount = work.get(top, ("hbm", 0.0))
    dur = per_op[top] * 1e-3
    if kind == "tensor":
        ach = amount / dur / 1e12
        return {"kernel": top, "bound": "tensor", "achieved": ach, "peak": pk["tensor"], "unit": "TFLOP/s",
                "frac": ach / pk["tensor"], "traffic": None, "ms": per_op[top], "share_of_step": share(per_op[top]),
                "note": "%s; peak = measured dense bf16 (sustained), %s" % (MATH_NOTE.get(args.precision, args.precision),
                                                                          pk["src"])}
    ach = amount / dur / 1e9
    return {"kernel": top, "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
            "traffic": None, "ms": per_op[top], "share_of_step": share(per_op[top]), "note": pk["src"]}


def make_batch(B, N, C, device, gen):
    # region features are stored as bf16 shards (engine.pack_feature_shard): the synthetic values are rounded to bf16
    # once, so the device-resident arm (fp32 copies of these values) and the host arm (bf16 over PCIe) see the same data
    v = torch.relu(torch.randn(B, N, D, generator=gen)).to(torch.bfloat16).to(torch.float32)
    q = 0.1 * torch.relu(torch.randn(B, Q, generator=gen))
    a = torch.zeros(B, C)
    cls = torch.randint(0, C, (B, 3), generator=gen)
    for k, mass in enumerate((0.6, 0.3, 0.1)):
        a.scatter_add_(1, cls[:, k:k + 1], torch.full((B, 1), mass))
    return v, q, a


# ----------------------------------------------------------------------------------------- CPU oracle
def cpu_oracle_rate(model, Bc, N, C, steps, warmup, train):
    """samples/s of the restated reference on the host cores (torch CPU, all threads)."""
    from oracle import reasoning_core as rc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = rc.synth_state_dict(model, C, seed=10, num_regions=N)
    gen = torch.Generator().manual_seed(1234)
    v, q, a = make_batch(Bc, N, C, "cpu", gen)
    drop = rc.torch_drop if train else rc.no_drop
    for _ in range(warmup):
        rc.step(model, sd, v, q, a, drop=drop, num_regions=N)
    t0 = time.perf_counter()
    for _ in range(steps):
        rc.step(model, sd, v, q, a, drop=drop, num_regions=N)
    dt = time.perf_counter() - t0
    return Bc * steps / dt, dt / steps, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    C = NUM_ANS[args.model]
    # The reference arm steps the SAME batch as the GPU arm unless the host cannot hold it: the reference form
    # materialises [B,N,N,2048] (CoR2) / [B,N,N*310] (ODA) tensors plus autograd copies, ~8 such tensors alive.
    if not args.cpu_batch:
        per_sample = 8 * 4 * args.regions * args.regions * (2048 if args.model == "CoR2" else 310)
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:
            avail = 32 << 30
        args.cpu_batch = max(2, min(args.batch, int(0.5 * avail / per_sample)))
    rate, per_step, cores = cpu_oracle_rate(args.model, args.cpu_batch, args.regions, C, args.steps,
                                            min(args.warmup, 2), not args.eval_mode)
    sample = "%s fwd+KLD+bwd, reference-form oracle (materialised pairwise/compound tensors), batch %d x %d regions, " \
             "%s mode, torch CPU %d threads, %d steps" % (args.model, args.cpu_batch, args.regions,
                                                          "eval" if args.eval_mode else "train", cores, args.steps)
    line = {
        "impl": "reference", "metric": "%s train samples/sec (fwd+bwd)" % args.model, "value": rate,
        "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 2),
        "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        # the CPU arm steps a bounded sample: batch --cpu-batch of the same workload (samples are independent, the metric
        # is per sample); "batch_note" keeps the GPU arm's batch visible
        "config": workload_config(args, C, batch=args.cpu_batch),
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, C, batch=None):
    return {"workload": "%s fwd+KLD-loss+bwd, batch %d/GPU x %d regions x 2048-d + 2400-d question, %d answers, "
                        "%s mode, %s" % (args.model, batch or args.batch, args.regions, C,
                                         "eval" if args.eval_mode else "train (Philox dropout p=0.5)",
                                         "stock config/CoR2.py chain (att1 -> compound -> att2)" if args.model == "CoR2"
                                         else "stock config/ODA.py"),
            "precision": args.precision, "parallelism": "dp%d" % args.gpus,
            "allreduce": getattr(args, "allreduce", None) or (
                "none (1 GPU)" if args.gpus == 1 else
                "bucketed all-reduce captured inside the step graph, overlapped with the backward"),
            "l2": "inputs rotate over 4 distinct batches and each step touches >0.6 GB of activations (> 126 MB L2)",
            "input_format": "region features stored as bf16 shards (values bf16-representable in both arms); e2e ships them "
                            "as bf16 over PCIe from pinned memory and widens them on the device"}


# ----------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import importlib
    import ctypes
    from vqa_playground_pytorch_b200 import _lib, ops
    from vqa_playground_pytorch_b200.parallel import DataParallelEngine
    cf = importlib.import_module("vqa_playground_pytorch_b200.config." + args.model)
    L = _lib.lib()
    C, B, N = NUM_ANS[args.model], args.batch, args.regions

    torch.manual_seed(10)                      # reference default init under manual_seed(10) (config/CoR2.py:244)
    model = cf.Model(None, C, num_regions=N, precision=args.precision).to(dev)
    model.train(not args.eval_mode)
    ops.manual_seed(1234 + rank)
    engine = DataParallelEngine(model, allreduce=args.transport)
    engine.broadcast_parameters()

    gen = torch.Generator().manual_seed(1234 + rank)
    host = [make_batch(B, N, C, "cpu", gen) for _ in range(4)]
    resident = [tuple(t.to(dev) for t in b) for b in host]

    def eager_step(v, q, a):
        logits = model({"v": v, "q_idxes": q})
        loss = ops.kld_loss(logits, a)
        loss.backward()
        engine.wait()
        return loss

    graphed = None
    if not args.no_graph:
        from vqa_playground_pytorch_b200.engine import GraphedStep
        v0, q0, a0 = resident[0]
        example = {"v": v0.clone(), "q_idxes": q0.clone(), "a": a0.clone()}
        overlap = world > 1 and not args.no_overlap
        try:
            graphed = GraphedStep(model, example, engine, capture_collectives=overlap)
        except Exception as exc:            # capture of the NCCL calls refused: reduce after the replay instead
            if not overlap:
                raise
            print("graph capture with collectives failed (%s); falling back to reduce-after-replay" % exc, file=sys.stderr)
            torch.cuda.synchronize()
            overlap = False
            graphed = GraphedStep(model, example, engine, capture_collectives=False)
        what = ("libvqacore peer-memory all-reduce over NVLink (vqa_peer_allreduce_f32%s, no shared memory: runs next to "
                "the backward's GEMMs)" % (", NVLS in-switch reduction" if engine.nvls else "")
                if engine.transport == "peer" else "NCCL all-reduce")
        args.allreduce = ("bucketed %s captured inside the step graph, overlapped with the backward" % what if overlap
                          else "bucketed %s after the graph replay" % what) if world > 1 else "none (1 GPU)"

    def step(v, q, a):
        if graphed is None:
            return eager_step(v, q, a)
        return graphed({"v": v, "q_idxes": q, "a": a})

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident inputs
    for i in range(args.warmup):
        step(*resident[i % 4])
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = L.vqa_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(*resident[i % 4])
    e1.record()
    barrier()
    sampler.stop_flag = True
    launches = L.vqa_launch_count() - launches0
    if graphed is not None:
        launches = graphed.launches_per_replay * args.steps
    ms = e0.elapsed_time(e1)
    if engine.peer_error():
        raise RuntimeError("a peer all-reduce gave up waiting for another rank: the timed steps are invalid")
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    value = world * B * args.steps / (ms * 1e-3)

    # ---- e2e: the public API with PINNED HOST inputs: every step copies its own batch host->device (prefetched on a
    # copy stream while the previous step computes — engine.HostPrefetcher) and reads the loss back (.item()).
    from vqa_playground_pytorch_b200.engine import HostPrefetcher, pack_feature_shard
    host_samples = pack_feature_shard([{"v": b[0], "q_idxes": b[1], "a": b[2]} for b in host])     # pinned, v as bf16

    prefetcher = HostPrefetcher([], dev, widen_into=graphed.static if graphed is not None else None)

    def e2e_run(nsteps):
        last = None
        pf = prefetcher.reset([host_samples[i % 4] for i in range(nsteps)])   # same device staging buffers every pass
        for smp in pf:
            last = step(smp["v"], smp["q_idxes"], smp["a"]).item()
        return pf.bytes_per_batch, last

    e2e_run(min(args.warmup, 3))
    barrier()
    e0.record()
    h2d, _ = e2e_run(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = t.item()
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)

    # ---- per-op timing (same step, CUDA events around every plan op) -> roofline of the dominant op
    roof, breakdown, kernels = None, None, None
    nprof = min(args.steps, 10)
    if rank == 0:
        L.vqa_profile_begin()
    engine.defer = False
    for i in range(nprof):                 # every rank steps (the step holds the gradient all-reduce); eager launches
        eager_step(*resident[i % 4])
    torch.cuda.synchronize()
    if rank == 0:
        buf = ctypes.create_string_buffer(1 << 16)
        L.vqa_profile_end(buf, len(buf))
        per_op, per_kernel = {}, {}
        for item in buf.value.decode().split(";"):
            if item:
                name, rest = item.rsplit("=", 1)
                tot, cnt = rest.split("/")
                (per_kernel if name.startswith("k:") else per_op)[name] = float(tot) / int(cnt)
        total = sum(per_op.values())
        pk = peaks()
        breakdown = {k: round(v, 4) for k, v in sorted(per_op.items(), key=lambda kv: -kv[1])}
        kernels = {k[2:]: round(v, 4) for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1])[:24]}
        roof = kernel_roofline(per_kernel, per_op, total, pk, args, B, N, C)
    if world > 1:
        dist.barrier()
    if world > 1:
        dist.barrier()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        args.cpu_batch = args.cpu_batch or 64
        rate, per_step, cores = cpu_oracle_rate(args.model, args.cpu_batch, N, C, 3, 1, not args.eval_mode)
        cpu = {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": "oracle (restated reference, torch CPU, %d threads), %s fwd+KLD+bwd at batch %d x %d regions, "
                         "3 steps after 1 warm-up (%.2f s/step)" % (cores, args.model, args.cpu_batch, N, per_step)}

    if rank == 0:
        line = {
            "metric": "%s train samples/sec (fwd+bwd)" % args.model, "value": value, "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            # fp32-parity modes deliver fp32 results (<= 1e-4 of the fp32 reference) whatever their MMA operand type,
            # which config.precision names; the reduced modes are named by their operand type
            "dtype": {"bf16": "bf16", "tf32": "tf32"}.get(args.precision, "f32"), "data": "synthetic",
            "config": workload_config(args, C),
            "clocks": sampler.result(),
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": roof, "cpu_baseline": cpu, "per_op_ms": breakdown, "top_kernels_ms": kernels,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # Leave without NCCL's communicator teardown: with collectives captured in a CUDA graph
        # destroy_process_group() was seen to block until the launcher's timeout.  Everything is measured and
        # printed at this point; the barrier keeps the ranks together until rank 0 has flushed its line.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
