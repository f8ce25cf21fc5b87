"""Shared Model plumbing for config/CoR2.py and config/ODA.py (reference Model.forward contracts)."""
import torch
import torch.nn as nn

from .. import ops


class CoreModel(nn.Module):
    """Holds the parameter containers (same names/order as the reference so state_dict matches) and
    runs Model.forward as ONE call into libvqacore (ops.ModelCoreFn).

    Extra constructor arguments, all defaulting to the reference's behaviour:
      num_regions  N (reference hard-codes 36: config/CoR2.py:203, config/ODA.py:202,222)
      precision    GEMM arithmetic (include/vqacore.h VQA_MATH_*).  fp32-parity modes (<= 1e-4 vs the reference):
                   'bf16x3' (default: the large region-side GEMMs read bf16 hi + lo operand planes and issue three
                   bf16 tcgen05 MMAs per product, the small ones run 3xTF32), 'tf32x3' (fp32 operands, 3xTF32
                   everywhere), 'fp32' (CUDA-core FMA, bitwise reproducible).  Reduced precision (<= 2e-2):
                   'bf16' (one bf16 plane for the large GEMMs, one TF32 pass for the small ones), 'tf32'.
      seq2vec      question encoder module; default passes sample['q_idxes'] through as the
                   2400-d embedding (blocks.QuestionPassThrough)
    """
    MODEL = None

    def _finish_init(self, layer_names, precision):
        self.precision = precision
        # dropout call-site ids in forward-call order == oracle/reasoning_core.py *_LAYERS
        for i, name in enumerate(layer_names):
            mod = self.get_submodule(name)
            mod.layer_id = i
        for m in self.modules():
            if hasattr(m, "math"):
                m.math = precision
        self.alpha_dict = {}
        self.grad_sink = None          # set by parallel.DataParallelEngine
        self.fixed_seed = None         # tests: pin the Philox key of the next train-mode forward
        self.seed_device = None        # int64[1] CUDA tensor: device-resident Philox key (engine.GraphedStep)
        self.use_seed_device = False   # only True inside GraphedStep's body: eager forwards draw ops.next_seed()

    def core_parameters(self):
        """Parameters in the reference's state_dict order, seq2vec.* excluded."""
        return [p for n, p in self.named_parameters() if not n.startswith("seq2vec.")]

    def _run_core(self, sample):
        v = sample['v']
        if hasattr(self.seq2vec, "graph_capture"):
            self.seq2vec.graph_capture = bool(self.use_seed_device)
        q = self.seq2vec(sample['q_idxes'])
        if v.numel() % (self.num_regions * 2048) != 0:
            raise ValueError("sample['v'] with %d elements cannot be viewed as [-1, %d, 2048]" %
                             (v.numel(), self.num_regions))
        train = bool(self.training)
        seed = 0
        if train:
            seed = self.fixed_seed if self.fixed_seed is not None else ops.next_seed()
        return ops.ModelCoreFn.apply(self.MODEL, v, q, train, self.precision, seed,
                                     self.seed_device if (train and self.use_seed_device) else None, self.num_regions,
                                     self.num_classes, self.grad_sink, *self.core_parameters())

    def __call__(self, *input, **kwargs):
        return super().__call__(*input, **kwargs)
