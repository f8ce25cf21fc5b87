/*
 * vqacore.h — C ABI of libvqacore_sm100a.so: the B200 (sm_100a) reasoning core shared by the
 * reference's two models, ODA (object-difference attention) and CoR2 (chain of reasoning).
 *
 * The reference (bupt-cist/vqa-playground-pytorch) is pure Python over ATen and has no FFI
 * layer; the boundary a maintainer binds is therefore this header, one entry point per
 * reference building block on the hot path.  Each declaration cites the
 * reference code it replaces (paths relative to the reference root).
 *
 * Conventions (all entry points):
 *   - plain C: POD structs, device pointers, sizes; no torch / C++ types cross the boundary;
 *   - the CALLER owns every buffer (inputs, outputs, workspace, gradient buffers); the library
 *     never allocates, frees or retains device memory past return;
 *   - all work is enqueued on the stream passed in (a cudaStream_t cast to void*); no hidden
 *     synchronisation, no default-stream use, safe to capture in a CUDA graph;
 *   - return 0 on success, a negative VQA_E* code otherwise; vqa_last_error() gives a
 *     thread-local message; no C++ exception crosses the boundary;
 *   - tensors are row-major fp32 unless a struct says otherwise; "ld*" are row strides in
 *     elements;
 *   - there is NO CPU fallback: without an sm_100 device every compute call fails with
 *     VQA_ENODEVICE / a CUDA launch error.
 *
 * Dropout (train mode).  The reference calls F.dropout(p) on the INPUT of every MyLinear /
 * MyConv1d (config/CoR2.py:77-78,115-116; config/ODA.py:94-95,128-129).  Here the mask is
 * never stored: it is regenerated from a counter-based Philox4x32-10 stream.  One Philox call covers 16
 * consecutive elements, one byte each:
 *     out  = Philox4x32-10(key = seed, ctr = (idx>>4 lo, idx>>4 hi, layer, 0))         (4 x 32 bits)
 *     byte = (out[(idx>>2)&3] >> 8*(idx&3)) & 0xFF;    keep(seed, layer, idx) = byte >= floor(p*256)
 * idx = row-major linear index of the element in the logical tensor the reference drops, scale 1/(1-p)
 * (p = 0.5, the only rate the reference uses, is represented exactly).  oracle/philox.py is the CPU twin used
 * by the parity tests.
 */
#ifndef VQACORE_H_
#define VQACORE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQA_ABI_VERSION 2

enum {
  VQA_OK = 0,
  VQA_EINVAL = -1,      /* bad argument (shape, null pointer, unsupported combination) */
  VQA_ECUDA = -2,       /* CUDA runtime / launch error, see vqa_last_error() */
  VQA_ENODEVICE = -3,   /* no sm_100 device */
  VQA_EWORKSPACE = -4   /* workspace too small */
};

enum { VQA_ACT_NONE = 0, VQA_ACT_RELU = 1, VQA_ACT_SIGMOID = 2, VQA_ACT_TANH = 3 /* GRU candidate state only */ };

/* GEMM arithmetic.  Every mode but FP32_SIMT runs on the tcgen05 tensor cores with fp32 accumulation in
 * TMEM and NEVER falls back to the CUDA-core GEMM: what a mode cannot run is VQA_EINVAL.
 *   FP32_SIMT  fp32 FMA on CUDA cores (exact fp32 arithmetic, bitwise reproducible)
 *   TF32X3     fp32 operands, error-compensated 3xTF32 (x = hi + lo, three kind::tf32 MMAs): fp32-parity, <= 1e-4
 *   TF32       fp32 operands, one TF32 pass: reduced precision, <= 2e-2
 *   BF16X3     fp32-parity on bf16 tensor cores: the large-M GEMMs (M >= 1024, the region-side contractions) read
 *              bf16 operand PLANES x = hi + lo (hi = bf16(x), lo = bf16(x - hi)) written by their producers and issue
 *              three kind::f16 MMAs per product (hi.hi + hi.lo + lo.hi); small-M GEMMs run TF32X3.  <= 1e-4
 *   BF16       reduced precision: large-M GEMMs on ONE bf16 plane, small-M GEMMs one TF32 pass.  <= 2e-2 */
enum { VQA_MATH_FP32_SIMT = 0, VQA_MATH_TF32X3 = 1, VQA_MATH_TF32 = 2, VQA_MATH_BF16 = 3, VQA_MATH_BF16X3 = 4 };

#define VQA_MAX_GROUPS 8
#define VQA_GLIMPSES 4

int vqa_abi_version(void);
const char* vqa_last_error(void);
/* 0 when an sm_100 device is current, VQA_ENODEVICE otherwise. */
int vqa_device_check(void);
/* sizeof() of a parameter struct by name ("vqa_linear_fwd_params", ...); 0 if unknown. Lets a
 * binding written in another language verify its mirror of the layouts below. */
size_t vqa_sizeof(const char* struct_name);
/* Number of kernels this library has launched in this process (all threads). */
unsigned long long vqa_launch_count(void);
/* Optional per-op timing of the whole-model plans: between begin and end every op of vqa_*_fwd/bwd is
 * bracketed by CUDA events on the caller's stream; end() synchronises them and writes
 * "op=total_ms/count;..." into buf. Used by bench.py for the roofline of the dominant kernel. */
int vqa_profile_begin(void);
int vqa_profile_end(char* buf, size_t cap);

typedef struct {
  float p;            /* drop probability; 0 disables */
  uint32_t layer;     /* Philox counter word 2 */
  uint64_t seed;      /* Philox key; change it every step */
  const uint64_t* seed_dev;  /* optional: when non-NULL the key is READ FROM DEVICE MEMORY at kernel run time and
                                `seed` is ignored, so a captured CUDA graph draws a fresh mask on every replay
                                (advance it with vqa_seed_advance inside the graph) */
} vqa_dropout;
/* *seed_dev += 1, enqueued on the stream. */
int vqa_seed_advance(uint64_t* seed_dev, void* stream);

/* Padded weight copies for the tensor-core paths.  TMA needs 16-byte row strides, which K = 310 / 510 weights do
 * not have; vqa_pack_weights copies each segment src [rows, K] into dst [rows_pad, roundup(K,4)] (zero filled) in
 * ONE launch, so a training step packs every such weight once and both its forward and backward reuse it
 * (Wp / W1p / W2p fields below).  Without them the ops pack into their own workspace on every call. */
typedef struct {
  const float* src; float* dst;
  int64_t rows, rows_pad, K;
} vqa_pack_segment;
#define VQA_MAX_PACK_SEGMENTS 24
int vqa_pack_weights(const vqa_pack_segment* segs, int nsegs, void* stream);

/* Packed dropout keep-bits: bit (i & 7) of out[i >> 3] = keep(seed, layer, i) for i < n (n rounded up to 16).
 * A cache of the Philox contract above for large inputs that several kernels drop with the SAME mask (the
 * forward GEMM, the wgrad GEMM and the dgrad epilogue of one layer): 1 bit per element instead of one Philox
 * call per 16 elements in every one of them.  out must hold (n + 15) / 16 * 2 bytes. */
int vqa_dropout_bits(float p, uint64_t seed, const uint64_t* seed_dev, uint32_t layer, uint64_t n, uint8_t* out,
                     void* stream);
/* The same for several dropout sites of one step in ONE launch (a plan calls it once at its head). */
typedef struct { uint32_t layer; uint64_t n; uint8_t* out; } vqa_bits_segment;
#define VQA_MAX_BITS_SEGMENTS 32
int vqa_dropout_bits_batch(float p, uint64_t seed, const uint64_t* seed_dev, const vqa_bits_segment* segs, int nsegs,
                           void* stream);

/* ------------------------------------------------------------------------------------------
 * Grouped linear:  Y_g = act( dropout_g(X_g) . W_g^T + b_g ),  g < groups.
 * Replaces MyLinear.forward (config/CoR2.py:106-119 == config/ODA.py:123-136) and
 * MyConv1d.forward with kernel_size 1 (config/CoR2.py:72-88 == config/ODA.py:89-105; the two
 * transposes vanish: a k=1 conv over [B,N,Cin] is this GEMM with M = B*N), and putils.Linear
 * (putils/__init__.py:25-30).  Groups share M,K,N,act,p and differ in pointers and dropout
 * layer; they run in one launch (e.g. the four 2400->310 question projections of CoR2,
 * the four glimpse linears of MyATT, config/CoR2.py:147-152).
 * dropout index of X_g[m,k] is  drop_index_base[g] + m*K + k.
 */
typedef struct {
  int groups;
  int64_t M, K, N;
  int act;
  int math;
  float p;
  uint64_t seed;
  const uint64_t* seed_dev;                  /* optional device-resident key, see vqa_dropout */
  const float* X[VQA_MAX_GROUPS];  int64_t ldx[VQA_MAX_GROUPS];
  const float* W[VQA_MAX_GROUPS];            /* [N,K] row-major (nn.Linear / Conv1d k=1 layout) */
  const float* b[VQA_MAX_GROUPS];            /* [N] or NULL */
  float* Y[VQA_MAX_GROUPS];        int64_t ldy[VQA_MAX_GROUPS];
  uint32_t layer[VQA_MAX_GROUPS];
  uint64_t drop_index_base[VQA_MAX_GROUPS];
  const uint8_t* drop_bits[VQA_MAX_GROUPS];  /* optional: packed keep-bits of X_g (bit i = element i) from
                                                vqa_dropout_bits[_batch](p, seed, layer[g], M*K); used when
                                                drop_index_base[g] == 0; the buffer needs 2 readable bytes past
                                                its last one (rows that are not a multiple of 4 long) */
  const float* Wp[VQA_MAX_GROUPS];           /* optional: W_g re-laid by vqa_pack_weights() as [N, roundup(K,4)] */
  void* workspace;          /* >= vqa_linear_fwd_workspace_bytes(); may be NULL when that is 0 */
  size_t workspace_bytes;
} vqa_linear_fwd_params;
/* Scratch the tensor-core paths need (padded copies of operands whose row stride is not a multiple of 16
 * bytes, which TMA cannot address; the activation-gradient dZ in the backward). 0 for VQA_MATH_FP32_SIMT.
 * The query sees shapes only: it covers W and a contiguous X when K is not a multiple of 4.  An X_g whose BASE is
 * not 16-byte aligned or whose ldx is not a multiple of 4 although K is needs groups * (M * roundup(K,4) * 4 + 256)
 * bytes more.  A tensor-core math mode never falls back to the CUDA-core GEMM: a workspace too small for the repack
 * is VQA_EINVAL. */
size_t vqa_linear_fwd_workspace_bytes(int math, int groups, int64_t M, int64_t K, int64_t N);
size_t vqa_linear_bwd_workspace_bytes(int math, int groups, int64_t M, int64_t K, int64_t N);
int vqa_linear_fwd(const vqa_linear_fwd_params* p, void* stream);

/* Backward of the grouped linear.  dZ = dY (.) act'(Y);  dW_g (+)= dZ^T . dropout(X);
 * db_g (+)= colsum(dZ);  dX_g (+)= (dZ . W_g) (.) mask/(1-p)   (dX_g NULL: skipped, as for the
 * graph inputs v and q).  accumulate_w / accumulate_x choose "+=" over "=" so that weight
 * gradients can land directly in a flat data-parallel gradient buffer.
 * Autograd of F.linear / F.conv1d / F.dropout / relu / sigmoid in the reference.
 */
typedef struct {
  int groups;
  int64_t M, K, N;
  int act;
  int math;
  float p;
  uint64_t seed;
  const uint64_t* seed_dev;
  int accumulate_w;
  int accumulate_x;
  const float* X[VQA_MAX_GROUPS];  int64_t ldx[VQA_MAX_GROUPS];
  const float* W[VQA_MAX_GROUPS];
  const float* Y[VQA_MAX_GROUPS];  int64_t ldy[VQA_MAX_GROUPS];   /* forward output (for act') */
  const float* dY[VQA_MAX_GROUPS]; int64_t lddy[VQA_MAX_GROUPS];
  float* dW[VQA_MAX_GROUPS];
  float* db[VQA_MAX_GROUPS];
  float* dX[VQA_MAX_GROUPS];       int64_t lddx[VQA_MAX_GROUPS];
  uint32_t layer[VQA_MAX_GROUPS];
  uint64_t drop_index_base[VQA_MAX_GROUPS];
  const uint8_t* drop_bits[VQA_MAX_GROUPS];  /* optional, as in the forward */
  const float* Wp[VQA_MAX_GROUPS];           /* optional, as in the forward */
  void* workspace;          /* >= vqa_linear_bwd_workspace_bytes() */
  size_t workspace_bytes;
  /* Optional addend fused into the dX_0 store (groups == 1, K % 4 == 0): the gradient of a 4-glimpse attention
   * pooling over the same X (MyATT's bmatmul, config/CoR2.py:118-121):
   *   dX_0[m, :] += sum_g pool_alpha[m*4 + g] * pool_dpooled[(m / pool_regions)*4*K + g*K + :]
   * pool_alpha NULL: none. */
  const float* pool_alpha;
  const float* pool_dpooled;
  int64_t pool_regions;
} vqa_linear_bwd_params;
int vqa_linear_bwd(const vqa_linear_bwd_params* p, void* stream);

/* ------------------------------------------------------------------------------------------
 * Mutan (bilinear) fusion:  Y[m,:] = sum_{r<R} (X1[m,:].W1_r^T + b1_r) (.) H2_r[m / rows_per_h2, :]
 * with H2_r = X2.W2_r^T + b2_r  ([Mh,F], Mh = M / rows_per_h2).
 * Replaces MutanFusion.forward (putils/__init__.py:232-238) and its per-sample bmul loop
 * (putils/__init__.py:98-104): rows_per_h2 = N regions when x1 is [B,N,.] and x2 is [B,.]
 * (fusion_vq1/2, config/CoR2.py:214,219), 1 for fusion_final (config/CoR2.py:233,
 * config/ODA.py:236).
 * H1 (optional, [R,M,F]) and H2 ([R,Mh,F]) are caller-owned stashes used by the backward.
 */
typedef struct {
  int R;
  int64_t M, K1, K2, F;
  int64_t rows_per_h2;
  int math;
  const float* X1; int64_t ldx1;
  const float* X2; int64_t ldx2;
  const float* W1[VQA_MAX_GROUPS]; const float* b1[VQA_MAX_GROUPS];
  const float* W2[VQA_MAX_GROUPS]; const float* b2[VQA_MAX_GROUPS];
  float* H1;      /* [R,M,F] or NULL */
  float* H2;      /* [R,Mh,F] */
  float* Y; int64_t ldy;
  const float* W1p;         /* optional: the R W1 matrices stacked by vqa_pack_weights as [R*roundup(F,32), roundup(K1,4)] */
  const float* W2p;         /* optional: likewise for W2 */
  void* workspace;          /* >= vqa_mutan_workspace_bytes(.., bwd=0); 256-byte aligned */
  size_t workspace_bytes;
} vqa_mutan_fwd_params;
/* Scratch of the tensor-core Mutan paths (padded stacked weights; dH1/dH2 in the backward); 0 for FP32_SIMT. */
size_t vqa_mutan_workspace_bytes(int math, int R, int64_t M, int64_t rows_per_h2, int64_t K1, int64_t K2, int64_t F,
                                 int bwd);
int vqa_mutan_fwd(const vqa_mutan_fwd_params* p, void* stream);

typedef struct {
  int R;
  int64_t M, K1, K2, F;
  int64_t rows_per_h2;
  int math;
  int accumulate_w;
  int accumulate_x1;
  int accumulate_x2;
  const float* X1; int64_t ldx1;
  const float* X2; int64_t ldx2;
  const float* W1[VQA_MAX_GROUPS];
  const float* W2[VQA_MAX_GROUPS];
  const float* H1;   /* [R,M,F] from forward */
  const float* H2;   /* [R,Mh,F] from forward */
  const float* dY; int64_t lddy;
  float* dH2;        /* workspace [R,Mh,F] */
  float* dW1[VQA_MAX_GROUPS]; float* db1[VQA_MAX_GROUPS];
  float* dW2[VQA_MAX_GROUPS]; float* db2[VQA_MAX_GROUPS];
  float* dX1; int64_t lddx1;     /* NULL: skipped */
  float* dX2; int64_t lddx2;     /* NULL: skipped */
  const float* W1p;         /* optional, as in the forward */
  const float* W2p;
  void* workspace;          /* >= vqa_mutan_workspace_bytes(.., bwd=1); 256-byte aligned */
  size_t workspace_bytes;
} vqa_mutan_bwd_params;
int vqa_mutan_bwd(const vqa_mutan_bwd_params* p, void* stream);

/* ------------------------------------------------------------------------------------------
 * Region softmax attention + pooling — the core of MyATT.forward
 * (config/CoR2.py:137-154 == config/ODA.py:154-171):
 *   z[b,i,g]   = sum_c Wc[g,c] * dropout(fuse)[b,i,c] + bc[g]        conv_att (Conv1d Ff->G, k=1)
 *   alpha[b,:,g] = softmax over regions i                            (af='softmax', dim=1)
 *   pooled[b,g,:] = sum_i alpha[b,i,g] * x[b,i,:]                    bmatmul, putils/__init__.py:89-95
 * The G glimpse linears that follow are a grouped vqa_linear_fwd on pooled.
 * dropout index of fuse[b,i,c] is (b*N+i)*Ff + c.
 */
typedef struct {
  int64_t B, N, Ff, D;
  vqa_dropout drop;
  const float* fuse;     /* [B,N,Ff] */
  const float* Wc;       /* [G,Ff] (Conv1d weight [G,Ff,1]) */
  const float* bc;       /* [G] */
  const float* x;        /* [B,N,D] */
  float* alpha;          /* [B,N,G] */
  float* pooled;         /* [B,G,D] */
  const uint8_t* drop_bits;  /* optional: keep-bits of fuse from vqa_dropout_bits(p, seed, layer, B*N*Ff) — the same
                                mask as the Philox contract, shared with the backward instead of being regenerated */
} vqa_region_softmax_pool_fwd_params;
int vqa_region_softmax_pool_fwd(const vqa_region_softmax_pool_fwd_params* p, void* stream);

/* Backward (SURVEY.md §8a "K-pool"):
 *   dalpha[b,i,g] = <dpooled[b,g,:], x[b,i,:]> + (g==0 ? dalpha0_ext[b] : 0) + dalpha_ext[b,i,g]
 *   dz = alpha (.) (dalpha - sum_i alpha*dalpha)
 *   dWc (+)= sum_{b,i} dz[b,i,g]*dropout(fuse)[b,i,c];  dbc (+)= sum dz
 *   dfuse[b,i,c] = (sum_g dz[b,i,g] Wc[g,c]) * mask/(1-p)
 *   dx[b,i,:] (+)= sum_g alpha[b,i,g] dpooled[b,g,:]      (dx NULL: x is a graph input)
 * dalpha0_ext ([B] or NULL) carries CoR2's d(s)/d(alpha) term from vqa_cor_compound_bwd; dalpha_ext ([B,N,G] or
 * NULL) is a general incoming gradient of alpha (a caller that uses the attention weights downstream, as the
 * reference's (alpha1[0] * v2_cat).sum(1) does, config/CoR2.py:216).
 */
typedef struct {
  int64_t B, N, Ff, D;
  vqa_dropout drop;
  int accumulate_w;
  int accumulate_x;
  const float* fuse; const float* Wc;
  const float* x; const float* alpha;
  const float* dpooled;        /* [B,G,D] */
  const float* dalpha0_ext;    /* [B] or NULL */
  float* dalpha;               /* workspace [B,N,G] */
  float* dz;                   /* workspace [B,N,G]; holds dz on return */
  float* dWc; float* dbc;
  float* dfuse;                /* [B,N,Ff] or NULL */
  float* dx;                   /* [B,N,D] or NULL */
  const uint8_t* drop_bits;    /* optional, as in the forward */
  const float* dalpha_ext;     /* [B,N,G] or NULL */
} vqa_region_softmax_pool_bwd_params;
int vqa_region_softmax_pool_bwd(const vqa_region_softmax_pool_bwd_params* p, void* stream);

/* ------------------------------------------------------------------------------------------
 * CoR2 compound objects.  Replaces decare_cat + the alpha-weighted sum
 * (config/CoR2.py:191-199, :215-216), which materialise [B,N,N,D]:
 *   v2_cat[b,i,j,:] = v[b,i,:]*g1[b,:] + v[b,j,:]*g2[b,:];  v2[b,j,:] = sum_i alpha1_0[b,i] v2_cat[b,i,j,:]
 * collapsed exactly to   v2[b,j,:] = vt[b,:]*g1[b,:] + s[b]*v[b,j,:]*g2[b,:],
 *   vt = sum_i alpha_i v_i  (= pooled[:,0,:] of att1),  s = sum_i alpha_i  (kept, not assumed 1).
 */
typedef struct {
  int64_t B, N, D;
  const float* x;          /* [B,N,D] block2 */
  const float* pooled;     /* [B,G,D]; glimpse 0 is vt */
  const float* alpha;      /* [B,N,G]; glimpse 0 gives s */
  const float* g1;         /* [B,D] */
  const float* g2;         /* [B,D] */
  float* v2;               /* [B,N,D] */
  /* Optional (bf16 math modes): also write the bf16 operand planes of dropout(v2) that compress_v2's GEMMs read
   * (VQA_MATH_BF16X3: 2 planes hi, lo; VQA_MATH_BF16: 1), [planes][B*N][D], planes v2_plane_stride elements apart;
   * keep = bit ((b*N+j)*D + c) of v2_keep_bits (NULL: all kept, scale 1).  Saves a second pass over v2. */
  void* v2_planes;         /* bf16, or NULL */
  int v2_nplanes;
  int64_t v2_plane_stride;
  const uint8_t* v2_keep_bits;
  float v2_keep_scale;
} vqa_cor_compound_fwd_params;
int vqa_cor_compound_fwd(const vqa_cor_compound_fwd_params* p, void* stream);

/* Backward: Dbar = sum_j dv2[j,:];  dg1 = vt*Dbar;  dpooled[b,0,:] += g1*Dbar;
 *   dg2 = s * sum_j v[j,:]*dv2[j,:];  dalpha0_ext[b] = sum_j <v[j,:]*g2, dv2[j,:]>. */
typedef struct {
  int64_t B, N, D;
  const float* x; const float* pooled; const float* alpha;
  const float* g1; const float* g2;
  const float* dv2;        /* [B,N,D] */
  float* dg1; float* dg2;  /* [B,D] */
  float* dpooled;          /* [B,G,D]; glimpse-0 slice is ACCUMULATED into */
  float* dalpha0_ext;      /* [B] */
  /* Optional: dv2 is handed over RAW by the producer of its largest term (compress_v2's dgrad GEMM, G = dZ.W) and
   * finished while it is read here, instead of in that GEMM's epilogue (one pass less over [B,N,D]):
   *   dv2_eff[b,j,c] = keep(b,j,c) * dv2_keep_scale * dv2[b,j,c] + sum_g dv2_pool_alpha[b,j,g] * dv2_pool_dpooled[b,g,c]
   * keep = bit ((b*N+j)*D + c) of dv2_keep_bits (compress_v2's input-dropout mask, vqa_dropout_bits; NULL: all kept),
   * the second term is the gradient of att2's pooling over the same v2 (NULL alpha: none). */
  const uint8_t* dv2_keep_bits;
  float dv2_keep_scale;
  const float* dv2_pool_alpha;     /* [B,N,G] */
  const float* dv2_pool_dpooled;   /* [B,G,D] */
} vqa_cor_compound_bwd_params;
int vqa_cor_compound_bwd(const vqa_cor_compound_bwd_params* p, void* stream);

/* ------------------------------------------------------------------------------------------
 * ODA object-difference attention logits + region softmax + pooling.  Replaces the 36x36
 * Python pair loop, the stack/transpose copy and conv_att over N*H channels
 * (config/ODA.py:216-226 with fuse_dim = N*310, :192):
 *   z[b,i,g] = bc[g] + sum_{j,k} W[g,j*H+k] * dropout(vq)[b,i,j*H+k],  vq = (vl[b,i,k]-vl[b,j,k])*ql[b,k]
 * The [B,N,N*H] tensor is never formed.  train = 0: factorised form
 * (z = sum_k ql*vl[i]*Wsum[k] + const_i, Wsum = sum_j W[g,j,:]); train = 1: every (i,j,k) term is
 * produced in registers with its keep flag (Philox index ((b*N+i)*N+j)*H+k).  Train mode needs a scratch buffer of
 * vqa_oda_pair_attn_workspace_bytes(B,N,H): the 1-bit-per-element keep cache of the dropped tensor
 * (vqa_dropout_bits layout, drawn ONCE per step) followed by the forward's partial logits.  The forward fills it,
 * the backward reads the keep bits back: pass the SAME buffer to both calls.  Train mode handles N <= 144 regions
 * (a CTA owns all regions of a feature range; the eval path has no such limit).
 */
size_t vqa_oda_pair_attn_workspace_bytes(int64_t B, int64_t N, int64_t H);
typedef struct {
  int64_t B, N, H, D;
  int train;
  vqa_dropout drop;
  const float* vl;       /* [B,N,H] */
  const float* ql;       /* [B,H] */
  const float* W;        /* [G,N*H] */
  const float* bc;       /* [G] */
  const float* x;        /* [B,N,D] */
  float* wsum;           /* workspace [G,H] (eval) */
  float* alpha;          /* [B,N,G] */
  float* pooled;         /* [B,G,D] */
  void* workspace;       /* train: vqa_oda_pair_attn_workspace_bytes(B,N,H) bytes, 256-byte aligned */
  size_t workspace_bytes;
  int keep_bits_ready;   /* nonzero: the keep bits at the head of the workspace are already drawn (a whole-model plan
                            makes them in its batched vqa_dropout_bits_batch launch); zero: this call draws them */
} vqa_oda_pair_attn_fwd_params;
int vqa_oda_pair_attn_fwd(const vqa_oda_pair_attn_fwd_params* p, void* stream);

typedef struct {
  int64_t B, N, H, D;
  int train;
  vqa_dropout drop;
  int accumulate_w;
  const float* vl; const float* ql; const float* W;
  const float* x; const float* alpha;
  const float* wsum;          /* from forward (eval) */
  const float* dpooled;       /* [B,G,D] */
  float* dalpha;              /* workspace [B,N,G] */
  float* dz;                  /* workspace [B,N,G]; holds dz on return */
  float* dwsum;               /* workspace [G,H] (eval) */
  float* dW; float* dbc;
  float* dvl;                 /* [B,N,H] */
  float* dql;                 /* [B,H] */
  const void* workspace;      /* train: the buffer the forward filled */
  size_t workspace_bytes;
} vqa_oda_pair_attn_bwd_params;
int vqa_oda_pair_attn_bwd(const vqa_oda_pair_attn_bwd_params* p, void* stream);

/* ------------------------------------------------------------------------------------------
 * Loss: KLDivLoss(size_average=False)(log_softmax(logits,1), target)  (train.py:536-544), fused
 * with its gradient:  loss_rows[b] = sum_c a*(log a - log_softmax(x)_c);
 * dlogits = grad_scale * (softmax(x) * sum_c a - a).  The scalar loss is sum(loss_rows).
 */
typedef struct {
  int64_t B, C;
  float grad_scale;
  const float* logits;   /* [B,C] */
  const float* target;   /* [B,C] */
  float* loss_rows;      /* [B] */
  float* dlogits;        /* [B,C] or NULL */
} vqa_kld_logsoftmax_params;
int vqa_kld_logsoftmax_fwd_bwd(const vqa_kld_logsoftmax_params* p, void* stream);
/* Input pipeline: region features stored and shipped as bf16 (pre-packed shards: half the host->device bytes of the
 * reference's fp32 h5 features, datasets.py:517-549, :905-970) are widened to the fp32 tensor the plans read.  Exact. */
int vqa_cast_bf16_f32(int64_t n, const void* src_bf16, float* dst, void* stream);
/* ------------------------------------------------------------------------------------------
 * Gradient all-reduce over NVLink peer memory: the ONE collective of data-parallel training (replaces the
 * gradient gather of nn.DataParallel, train.py:517).  buffers[r] / signals[r] are rank r's gradient buffer and signal
 * buffer AS MAPPED IN THE CALLING PROCESS (symmetric allocations exchanged once by the host side, e.g. CUDA IPC or
 * torch symmetric memory); every rank calls this with the same offset / count in the same order on its own stream.
 * In place: on return (stream order) buffers[rank][offset .. offset+count) holds the sum over the ranks, bit-identical
 * on every rank (each element is summed by one rank in rank order and broadcast).  The signal buffer
 * (vqa_peer_allreduce_signal_bytes() bytes per rank) must be zeroed once before the first call.  The kernel takes no
 * shared memory: it runs next to the persistent GEMMs of the backward instead of displacing them.
 * spin_limit_ms > 0: a rank that waits longer than that for a peer sets word [last] of its signal buffer to 1 and
 * gives up instead of hanging (the result is then invalid); 0 = wait forever. */
#define VQA_AR_MAX_WORLD 8
#define VQA_AR_MAX_CTAS 160
typedef struct {
  int world, rank;
  void* buffers[VQA_AR_MAX_WORLD];
  void* signals[VQA_AR_MAX_WORLD];
  int64_t offset, count;     /* in floats, multiples of 4 */
  int max_ctas;              /* 0 = default (160 CTAs of 128 threads); every rank must pass the same value */
  int spin_limit_ms;
  void* multicast;           /* optional: the MULTICAST mapping of the same symmetric buffer (NVLS).  When non-NULL the
                                slice is reduced inside the NVSwitch (multimem.ld_reduce) and broadcast by it
                                (multimem.st): half the NVLink bytes of the peer loads + stores.  All ranks or none. */
  int cta_threads;           /* 0 / 128: light CTAs (default); 256: wide CTAs, more loads in flight.  Same on all ranks. */
} vqa_peer_allreduce_params;
size_t vqa_peer_allreduce_signal_bytes(void);
int vqa_peer_allreduce_f32(const vqa_peer_allreduce_params* p, void* stream);

/* ------------------------------------------------------------------------------------------
 * SkipThoughts question encoder (putils/__init__.py:878-985): nn.Embedding(V, 620, padding_idx=0) ->
 * BayesianGRU(620 -> 2400) -> the hidden state at each question's last non-PAD token.  The dense contractions are the
 * grouped linears above (the three input projections of ALL time steps in one launch, the three recurrent projections
 * of a step in one launch, one wgrad over all steps at the end); these entry points are what sits between them.
 *
 * SequentialDropout (putils/__init__.py:503-539): ONE mask per sequence and call site, shared by every time step.
 * masks[m][b][f] = keep ? 1/(1-p) : 0 with keep from the Philox contract above, layer = layer0 + m, element index
 * b*dim + f.  p = 0 writes ones. */
int vqa_seq_dropout_masks(float p, uint64_t seed, const uint64_t* seed_dev, uint32_t layer0, int64_t B, int64_t dim,
                          int nmasks, float* out /* [nmasks][B][dim] */, void* stream);
/* out[g][(t*B + b)][f] = emb[idx[b][t]][f] * masks[g][b][f], g < 3 (drop_ir / drop_ii / drop_in of
 * BayesianGRUCell.forward :624-626); rows are TIME-major so that a step's slice is contiguous.  masks NULL = eval. */
int vqa_gru_embed_fwd(int64_t B, int64_t T, int64_t I, const int64_t* idx /* [B][T] */, const float* emb /* [V][I] */,
                      const float* masks /* [3][B][I] or NULL */, float* out /* [3][T*B][I] */, void* stream);
/* demb[idx[b][t]][f] += sum_g dX[g][(t*B+b)][f] * masks[g][b][f]; the PAD row 0 receives nothing (padding_idx=0). */
int vqa_gru_embed_bwd(int64_t B, int64_t T, int64_t I, const int64_t* idx, const float* masks, const float* dX,
                      float* demb, void* stream);
/* One step of BayesianGRUCell.forward (:630-636) after its six projections:
 *   r = sigmoid(gi_r + gh_r), i = sigmoid(gi_i + gh_i), n = act(gi_n + r*gh_n), h = (1-i)*n + i*h_prev
 * plus the masked copies hm_g = h * hmask_g the next step's recurrent projections read (drop_hr/hi/hn).
 * gh and h_prev are NULL together at the first step (h_prev = 0, the hidden projections have no bias). */
typedef struct {
  int64_t B, H;
  int act;                 /* VQA_ACT_RELU | VQA_ACT_TANH (SkipThoughts(af=...), config/CoR2.py:166-167) */
  const float* gi[3];      /* [B,H] each: input projections r, i, n of this step */
  const float* gh[3];      /* [B,H] each: recurrent projections, or NULL */
  const float* h_prev;     /* [B,H] or NULL */
  const float* hmask[3];   /* [B,H] sequence-tied masks or NULL (eval) */
  float* h;                /* [B,H] */
  float* r; float* i; float* n;   /* [B,H] stash for the backward (NULL to skip) */
  float* hm[3];            /* [B,H] masked copies for the next step (NULL to skip) */
} vqa_gru_gate_fwd_params;
int vqa_gru_gate_fwd(const vqa_gru_gate_fwd_params* p, void* stream);
/* Backward of one step.  dh = dh_partial + sum_g dhm[g]*hmask[g] + (last_pos[b] == t ? dx_last[b] : 0);
 * outputs da[0..2] = gradients of the pre-activations (= of gi_r, gi_i, gi_n and of gh_r, gh_i), dgh_n (of gh_n) and
 * dh_partial_out = dh * i (the direct path to h_prev). */
typedef struct {
  int64_t B, H;
  int act;
  int64_t t;
  const float* dh_partial;   /* [B,H] or NULL */
  const float* dhm[3];       /* [B,H] gradients of the NEXT step's masked inputs, or NULL */
  const float* hmask[3];
  const float* dx_last;      /* [B,H] gradient of the encoder output, or NULL */
  const int64_t* last_pos;   /* [B] */
  const float* r; const float* i; const float* n;
  const float* gh_n;         /* [B,H] or NULL (first step) */
  const float* h_prev;       /* [B,H] or NULL */
  float* da[3];
  float* dgh_n;
  float* dh_partial_out;
} vqa_gru_gate_bwd_params;
int vqa_gru_gate_bwd(const vqa_gru_gate_bwd_params* p, void* stream);
/* last_pos[b] = (#tokens != 0 of question b) - 1, -1 wrapping to T-1 like `mask[i][lengths[i] - 1]` (:729-731);
 * out[b] = hs[last_pos[b]][b] for the time-major hidden states hs [T][B][H]. */
int vqa_gru_last_pos(int64_t B, int64_t T, const int64_t* idx, int64_t* last_pos, void* stream);
int vqa_gru_select_last(int64_t B, int64_t H, const float* hs, const int64_t* last_pos, float* out, void* stream);

/* Prediction tail of the eval loop (train.py:146-169, `output.data.cpu().max(1)`): pred[b] = index of the first maximum
 * of logits[b, :] (OpenEnded), or of the candidates mc_idx[b, 0..n_mc) (MultipleChoice, train.py:153-164: -1 entries are
 * padding; pred = -1 when a row has no candidate).  best ([B] or NULL) receives the winning logit.  The B x C logits
 * never leave the device: only B indices do. */
int vqa_argmax_rows(int64_t B, int64_t C, const float* logits, const int64_t* mc_idx, int64_t n_mc, int64_t* pred,
                    float* best, void* stream);
/* out[0] = sum(rows[0..n)) with a fixed reduction order; y = x * scale[0] with the scalar read on the device
 * (the incoming autograd gradient of the scalar loss) — the two pieces that keep the loss off ATen. */
int vqa_sum_rows(int64_t n, const float* rows, float* out, void* stream);
int vqa_scale_by_device_scalar(int64_t n, const float* x, const float* scale, float* y, void* stream);

/* ------------------------------------------------------------------------------------------
 * Whole-model plans: one call enqueues every kernel of Model.forward / its backward
 * (config/CoR2.py:201-237, config/ODA.py:200-240), nothing from ATen in between.
 * Parameters are passed as a table of pointers in the reference's state_dict order
 * (seq2vec.* excluded): CoR2 62 tensors, ODA 38.  `grads` uses the same order.
 */
#define VQA_COR2_NPARAMS 62
#define VQA_ODA_NPARAMS 38

typedef struct {
  int64_t B, N, C;           /* batch, regions, answers */
  int train;                 /* 1: dropout on (Philox, `seed`) */
  int math;
  uint64_t seed;
  const uint64_t* seed_dev;  /* optional device-resident Philox key (CUDA graphs), see vqa_dropout */
  const float* v;            /* [B,N,2048] */
  const float* q;            /* [B,2400] */
  const float* const* params;
  float* logits;             /* [B,C] */
  float* alpha1;             /* [B,N,G] */
  float* alpha2;             /* [B,N,G]  (ODA: unused) */
  float* v2;                 /* [B,N,2048] (ODA: unused) */
  void* workspace; size_t workspace_bytes;
} vqa_model_fwd_params;

#define VQA_MAX_GRAD_GROUPS 16
typedef struct {
  vqa_model_fwd_params fwd;  /* same values as the forward call (workspace holds its stash) */
  const float* dlogits;      /* [B,C] */
  float* const* grads;       /* table of gradient pointers, state_dict order */
  int accumulate;            /* 1: grads += , 0: grads = */
  float* grads_flat;         /* optional: one buffer that contains every gradient tensor (e.g. the data-parallel
                                flat buffer); with accumulate == 0 the plan zero-fills it ONCE and lets every
                                kernel accumulate, instead of one memset per tensor */
  size_t grads_flat_bytes;
  /* Optional: cudaEvent_t handles recorded as the plan finishes its gradient GROUPS (vqa_grad_groups below), on the
   * stream that produced the group, so that a data-parallel caller can all-reduce a bucket of the flat buffer
   * while the rest of the backward is still running.  NULL entries are skipped. */
  void* group_events[VQA_MAX_GRAD_GROUPS];
  float* dq;                 /* optional [B,2400]: gradient of the loss w.r.t. the question embedding, for a trainable
                                question encoder in front of the core (seq2vec, config/CoR2.py:205); NULL skips the
                                extra dgrads of the 2400->310 question projections */
} vqa_model_bwd_params;
/* ------------------------------------------------------------------------------------------
 * Optimizer step over the flat gradient buffer: global-norm clipping + Adam in two launches.
 * Replaces nn.utils.clip_grad_norm_(model.parameters(), 0.25) + torch.optim.Adam.step()
 * (train.py:82-86; optimizer at train.py:292: lr, betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad):
 *   clip  = min(1, max_norm / (||grads_flat||_2 + 1e-6))            (max_norm <= 0: no clipping)
 *   g     = grad * clip;  m += (g - m)(1 - beta1);  v = beta2 v + (1 - beta2) g^2
 *   param -= lr / (1 - beta1^step) * m / (sqrt(v) / sqrt(1 - beta2^step) + eps)
 * Parameters stay where they are (segs[i].param); gradients and both moment buffers are flat, segment i at
 * [offset, offset + numel).  Every element of [0, total) must belong to exactly one segment for the norm to be the
 * norm of the gradients (a data-parallel flat buffer is exactly that).
 */
typedef struct { float* param; int64_t offset; int64_t numel; } vqa_param_segment;
#define VQA_MAX_PARAM_SEGMENTS 64
#define VQA_CLIP_SCRATCH_FLOATS 1024
typedef struct {
  int nsegs;
  const vqa_param_segment* segs;   /* host array */
  float* grads_flat;               /* [total], 16-byte aligned */
  float* exp_avg;                  /* [total] first moment  (zero before step 1) */
  float* exp_avg_sq;               /* [total] second moment */
  int64_t total;
  float lr, beta1, beta2, eps;
  int64_t step;                    /* 1 for the first update */
  float max_norm;                  /* clip_grad_norm_'s max_norm; <= 0 disables clipping */
  int write_clipped_grads;         /* 1: grads_flat *= clip as clip_grad_norm_ does in place; 0: leave them */
  float* scratch;                  /* VQA_CLIP_SCRATCH_FLOATS floats of device memory when clipping: [0] holds ||g||^2 on
                                      return, the rest the per-block partials of its fixed-order reduction */
  /* Optional device-resident optimizer clock, for a step captured in a CUDA graph (every replay must be the same
   * launch): when step_dev is non-NULL the call first enqueues *step_dev += 1 (and *lr_dev *= lr_gamma when lr_dev is
   * non-NULL and lr_gamma > 0 — ExponentialLR stepped BEFORE the optimizer, train.py:75-76, :296), then the update
   * reads the step count from *step_dev and the learning rate from *lr_dev (lr above when lr_dev is NULL);
   * `step` is ignored. */
  int64_t* step_dev;
  double* lr_dev;
  double lr_gamma;
} vqa_clip_adam_params;
int vqa_clip_adam_step(const vqa_clip_adam_params* p, void* stream);

/* Gradient groups of a backward plan in the order they complete: writes group_of_param[i] (i in state_dict order,
 * n_params entries) and returns the number of groups (<= VQA_MAX_GRAD_GROUPS), or -1 for an unknown model
 * (0 = CoR2, 1 = ODA) / wrong n_params.  Pure host function (no GPU needed). */
int vqa_grad_groups(int model, int* group_of_param, int n_params);

size_t vqa_cor2_workspace_bytes(int64_t B, int64_t N, int64_t C);
int vqa_cor2_fwd(const vqa_model_fwd_params* p, void* stream);
int vqa_cor2_bwd(const vqa_model_bwd_params* p, void* stream);

size_t vqa_oda_workspace_bytes(int64_t B, int64_t N, int64_t C);
/* Introspection for tests: where a ReLU output of the forward plan lives inside the workspace.
 * model: 0 = CoR2, 1 = ODA; name: "compress_v", "compress_v2", "compress_q", "compress_q_1", "compress_q_2",
 * "linear_q", "glimpses" (the concatenated glimpse-linear outputs). */
int vqa_stash_info(int model, const char* name, int64_t B, int64_t N, int64_t C, size_t* offset_bytes,
                   int64_t* rows, int64_t* cols, int64_t* ld);
int vqa_oda_fwd(const vqa_model_fwd_params* p, void* stream);
int vqa_oda_bwd(const vqa_model_bwd_params* p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VQACORE_H_ */
