/*
 * m2d.h — C ABI of libm2d_b200.so: hand-written sm_100a kernels for the phase3
 * audio-to-dance WGAN-GP training step of clementabary/music2dance.
 *
 * The reference has no native layer: every operation below is, in the reference,
 * a PyTorch library call (ATen -> cuDNN/cuBLAS/MKL-DNN).  Each entry point cites
 * the reference call site(s) (relative to the reference root) it replaces.
 *
 * Conventions
 *   - plain C types only; device pointers are `float*` / `const float*`;
 *     `stream` is a cudaStream_t passed as void* (NULL = legacy default stream)
 *   - all work is enqueued on `stream`; no host synchronisation, no allocation:
 *     the caller (PyTorch caching allocator) owns every buffer and workspace
 *   - return 0 on success, negative m2d_status otherwise; m2d_last_error() gives
 *     a thread-local message.  No C++ exceptions cross the boundary.
 *   - activations are CHANNELS-LAST row matrices: element (batch b, row l, col c)
 *     lives at  ptr[b*bs + l*ld + c]   (bs = batch stride, ld = row stride, floats)
 *   - the library refuses to run on anything but compute capability 10.x.
 */
#ifndef M2D_H
#define M2D_H

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    M2D_OK = 0,
    M2D_ERR_BAD_ARG = -1,
    M2D_ERR_CUDA = -2,
    M2D_ERR_ARCH = -3,
    M2D_ERR_WORKSPACE = -4
} m2d_status;

enum { M2D_ACT_NONE = 0, M2D_ACT_RELU = 1, M2D_ACT_LEAKY = 2, M2D_ACT_TANH = 3 };
/* derivative masks: value v is multiplied by act'(.) evaluated from the stored
 * post-activation tensor `mask`:  RELU: mask>0 ? 1 : 0;  LEAKY: mask>0 ? 1 : 0.2;
 * TANH: 1 - mask^2 */
enum { M2D_MASK_NONE = 0, M2D_MASK_RELU = 1, M2D_MASK_LEAKY = 2, M2D_MASK_TANH = 3 };

const char* m2d_last_error(void);
int m2d_version(void);
/* 0 if device `dev` is sm_100-class, M2D_ERR_ARCH otherwise. */
int m2d_check_device(int dev);

/* Arithmetic of the GEMM family (m2d_rowconv, m2d_wgrad).  The reference runs these
 * contractions as fp32 cuDNN/cuBLAS calls (phase3/train.py:33-34 leaves TF32 at the
 * PyTorch default).  Process-wide setting, not thread-safe against in-flight calls.
 *   FP32   : CUDA-core FFMA kernels (fp32 products, fp32 accumulate)
 *   TF32   : tcgen05 tensor cores, operands rounded to TF32 (round-to-nearest), fp32 accumulate
 *   TF32X3 : tcgen05 tensor cores, 3xTF32 operand split (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo):
 *            fp32-grade results on the tensor pipe — the default
 *   TF32_BF16 : tcgen05 tensor cores, a_hi*b_hi as TF32 plus both cross terms as ONE BF16 contraction over
 *            [bf16(a_hi) | bf16(a_lo)] x [bf16(b_lo) | bf16(b_hi)] (kind::f16, K = 16 per instruction): 2/3 of the
 *            tensor-pipe work of TF32X3 at fp32-grade operand error (6-7e-7, tools/split_precision_study.py);
 *            row convolutions only — weight gradients stay TF32X3.  Opt-in; weights must be re-packed
 *            (m2d_pack_batch) after switching to or from this mode: the lo plane of w_tiled changes format
 * Shapes the tensor-core kernels do not cover (N < 8, tiny problems, weight gradients with
 * Cout or Cc not a multiple of 4) run on the FP32 kernels in every mode. */
enum { M2D_GEMM_FP32 = 0, M2D_GEMM_TF32 = 1, M2D_GEMM_TF32_BF16 = 2, M2D_GEMM_TF32X3 = 3 };
/* number of m2d_rowconv calls served by the TMA halo-tile kernel so far (diagnostics / tests) */
long long m2d_halo_launch_count(void);
/* ... of which by its persistent variant (one CTA per SM over a static tile list, double-buffered TMEM accumulators,
 * epilogue overlapped with the next tile's MMAs): launches with >= 148 tiles and no split-K */
long long m2d_halo_persist_launch_count(void);
int m2d_set_gemm_mode(int mode);
int m2d_get_gemm_mode(void);

/* ------------------------------------------------------------------------
 * Row-convolution GEMM: the one contraction that serves Conv1d forward,
 * Conv1d backward-data (per stride residue), the WGAN-GP tangent pass and
 * every Linear layer.
 *
 *   y[b,i,n] = epi( sum_{t<T} sum_{c<Cc}  x[b, i*sr + roff0 + t*droff, c] * w[n, t*Cc + c] )
 *
 * rows outside [0, x_rows) read as zero (= Conv1d zero padding).
 * Windowed mode (win_T > 0, Cc == 1): x is raw audio [nseq, win_seq_len]; batch
 * b = seq*win_T + f addresses window f of sequence seq, i.e. sample
 * f*win_stride - win_pad + r, zero outside the sequence — utils.py:329-353
 * (slice_audio_batch) fused into the first encoder convolution.
 *
 * epi(v): v += bias[n]; v = act(v); [v += add  if add_before_mask]; [y2 = v];
 *         v *= act'(mask); [v += add  otherwise]; y = v
 *
 * Replaces: nn.Conv1d.forward in phase3/archis/default.py:64-70,90-97,117-127,
 * 202-210,217,298-303,326-333; nn.Linear.forward :153,161,175-176,256-257,280-281;
 * conv/linear backward-data and the double-backward "ggO" pass reached through
 * losses.py:40-44 and phase3/train.py:215,236.
 * ---------------------------------------------------------------------- */
typedef struct {
    const float* x; long long x_bs; int x_ld; int x_rows;
    int nb;
    int win_T, win_stride, win_pad, win_seq_len;
    const float* w; int w_ld;
    const float* w_tiled;                /* optional pre-split, pre-tiled copy of w for the tensor-core kernels (written
                                            by m2d_pack_batch): the weight operand as the tensor core reads it from shared
                                            memory, one contiguous block per (N tile, tap, 32-channel block) so that a stage
                                            is ONE bulk copy.  Block ((nt*T + t)*ceil(Cc/32) + c) holds R = (N <= 64 ? 64 : 128)
                                            rows x 32 floats of w_hi = tf32_rn(w) followed by the same of w_lo =
                                            tf32_rn(w - w_hi), rows 128 bytes apart, 16-byte chunks XOR-swizzled with
                                            (row & 7) (SWIZZLE_128B K-major); entries beyond N / Cc are zero.  Cc == 1: the
                                            taps play the role of channels (T = 1).  NULL: operands are split on the fly */
    int N, T, Cc;
    int sr, roff0, droff;
    float* y; long long y_bs; int y_ld; int y_rows;
    float* y2;                           /* optional: value before the mask / post-add (same geometry as y) */
    const float* bias;
    int act;
    const float* mask; long long m_bs; int m_ld; int mask_mode;
    const float* add; long long a_bs; int a_ld; int add_before_mask;
    float* ws; long long ws_floats;      /* split-K workspace of the FP32 kernels (may be NULL); the tensor-core
                                            kernels reduce split-K partials over a thread-block cluster (DSMEM) */
} m2d_rowconv_args;
int m2d_rowconv(const m2d_rowconv_args* a, void* stream);

/* ------------------------------------------------------------------------
 * Weight gradient of a row convolution, written in the PyTorch parameter
 * layout (Cout, Cc, T):
 *   dw[co,c,t] = beta*dw[co,c,t] + scale * sum_{b,l} dy[b,l,co] * x[b, l*sr + roff0 + t*droff, c]
 * Two-stage deterministic split-K through `ws`.
 * Replaces: conv/linear backward-weight and the double-backward "gW" pass
 * (autograd of default.py layers; train.py:215,236; losses.py:40-44).
 * ---------------------------------------------------------------------- */
typedef struct {
    const float* dy; long long dy_bs; int dy_ld; int dy_rows;
    int nb;
    const float* x; long long x_bs; int x_ld; int x_rows;
    int win_T, win_stride, win_pad, win_seq_len;
    int Cout, T, Cc;
    int sr, roff0, droff;
    float* dw;
    int packed;                          /* 0: dw in the PyTorch layout (Cout, Cc, T);  1: dw[co, t*Cc + c] (coalesced
                                            stores; m2d_pack_batch kind M2D_UNPACK_GRAD turns it into the PyTorch layout) */
    float scale, beta;
    float* ws; long long ws_floats;
} m2d_wgrad_args;
int m2d_wgrad(const m2d_wgrad_args* a, void* stream);
/* minimum workspace (floats) m2d_wgrad needs for these shapes */
long long m2d_wgrad_min_ws(int Cout, int T, int Cc);

/* Re-layout of PyTorch Conv1d weights (Cout, Cin, k):
 *   fwd : wp[co, t*Cin + ci]            = w[co, ci, t]
 *   bwd : wd_rho[ci, q*Cout + co]       = w[co, ci, stride*q + rho],  rho = 0..stride-1,
 *         blocks stored back to back, block rho has T_rho = ceil((k-rho)/stride) taps. */
int m2d_pack_conv_fwd(const float* w, float* wp, int Cout, int Cin, int k, void* stream);
int m2d_pack_conv_bwd(const float* w, float* wd, int Cout, int Cin, int k, int stride, void* stream);

/* All re-layouts of one network in a single launch.  `table` is an array of n descriptors in
 * DEVICE memory (pointers are stable, so it is built once).  kind FWD / BWD as above; FULL_BWD is
 * the backward layout of a convolution whose kernel spans its whole input (fconv, l6, encoder heads),
 * used as a Linear over (tap, channel):  dst[(t*Cin + ci), co] = w[co, ci, t]. */
enum { M2D_PACK_FWD = 0, M2D_PACK_BWD = 1, M2D_PACK_FULL_BWD = 2,
       M2D_UNPACK_GRAD = 3 /* dst[co, ci, t] = w[co, t*Cin + ci]: packed weight gradient -> parameter layout */,
       M2D_PACK_BWD_MERGED = 4 /* all stride residues in one matrix [(r0, ci)][q'*Cout + co] (unified taps,
                                  `reserved` = conv padding): backward-data as one stride-1 row convolution
                                  whose N = stride*Cin columns are the fine rows of a coarse output row */ };
typedef struct {
    const float* w; float* dst;           /* dst may be NULL (no exact copy wanted) */
    float* dst_tiled;                     /* optional 3xTF32 split copy in the m2d_rowconv_args.w_tiled layout of the
                                             GEMM the kind defines (FWD: N = Cout, T = k, Cc = Cin, or T = 1, Cc = k when
                                             Cin == 1; FULL_BWD: N = k*Cin, T = 1, Cc = Cout; BWD: per residue N = Cin,
                                             T = T_rho, Cc = Cout, blocks back to back; BWD_MERGED: N = stride*Cin, T = Tm,
                                             Cc = Cout).  Padding is never written: allocate zeroed */
    int Cout, Cin, k, stride, kind, reserved;
} m2d_pack_desc;
int m2d_pack_batch(const m2d_pack_desc* table, int n, void* stream);

/* Backward-data of a Conv1d with ONE input channel (AudioDiscriminator.l1,
 * default.py:298): dx[b,i] = sum_{co,j} dy[b,(i+pad-j)/stride,co] * w[co,0,j]. */
int m2d_conv_dgrad_c1(const float* dy, int nb, int Lout, int Cout, const float* w, int k,
                      int stride, int pad, float* dx, int Lin, void* stream);

/* ------------------------------------------------------------------------
 * GRU (torch.nn.GRU, batch_first, h0 = 0, gates [r|z|n]) — default.py:349-355,
 * used at :19-20,35-38.  gi = W_ih x + b_ih is precomputed with m2d_rowconv.
 * Persistent kernel: one thread-block cluster per batch group, W_hh rows sharded
 * across the cluster's CTAs in shared memory, h exchanged through DSMEM.
 *   h_out[b,t,:H] (row stride ldh)   save[b,t,4H] = r|z|n|(W_hn h + b_hn)
 * ---------------------------------------------------------------------- */
/* forward implementation: 2 (default) = recurrent weights resident in REGISTERS, warp-shuffle gate
 * reductions, h pushed to the cluster with st.async + mbarrier transaction counts (one mbarrier wait per
 * step); 1 = weights in shared memory, cluster barrier per step (any H up to the smem limit). */
int m2d_set_gru_impl(int impl);
/* Sequences served by one thread-block cluster of the forward recurrence (impl 2): 0 = automatic (one sequence per
 * cluster up to batch 18: lowest latency), 1..8 = that many (fewer SMs busy for slightly longer steps — what the fused
 * trainer asks for, because its generator forwards run on a side stream next to the critic iterations). */
int m2d_set_gru_forward_batch_group(int bg);
int m2d_gru_forward(const float* gi, const float* w_hh, const float* b_hh,
                    float* h_out, int ldh, float* save, int B, int T, int H, void* stream);
/* BPTT: dh_out[b,t,:H] (row stride ldd) upstream gradient; writes dgi, dgh [B,T,3H]. */
int m2d_gru_backward(const float* dh_out, int ldd, const float* h_out, int ldh,
                     const float* save, const float* w_hh,
                     float* dgi, float* dgh, int B, int T, int H, void* stream);

/* ------------------------------------------------------------------------
 * BatchNorm1d (train mode, eps, momentum) over the rows of a channels-last matrix —
 * default.py:65,68,91,94,118-127,154,179-180,217.
 *   stats   : acc[0:C] += column sums, acc[C:2C] += column sums of squares (fp64)
 *   apply   : y = act(gamma*(x-mean)*rstd + beta); block 0 also writes mean/rstd
 *             to `mr` (2C floats) and updates running_mean/var (unbiased var);
 *             y == NULL: statistics / running-stat update only
 *   eval    : y = act(gamma*(x-rm)/sqrt(rv+eps)+beta)
 *   bwd_red : acc[0:C] += sum dyy, acc[C:2C] += sum dyy*xhat, dyy = dy*act'(y)
 *   bwd_app : dx = gamma*rstd*(dyy - acc0/M - xhat*acc1/M); also dgamma = acc1, dbeta = acc0
 * ---------------------------------------------------------------------- */
int m2d_colstats(const float* x, int ld, long long M, int C, double* acc, void* stream);
int m2d_bn_apply(const float* x, int ldx, float* y, int ldy, long long M, int C,
                 const double* acc, const float* gamma, const float* beta,
                 float* running_mean, float* running_var, float momentum, float eps,
                 float* mr, int act, void* stream);
/* The same over `groups` consecutive row blocks of M rows each, one launch: block z is normalised with its OWN batch
 * statistics (acc + z*2C; mr + z*2C when given) and the running statistics advance block by block in order — i.e.
 * `groups` train-mode forwards of the same module on `groups` batches (the generator forwards of several critic
 * iterations, train.py:193-196, evaluated as one batched pass; the generator's weights do not change in between). */
int m2d_colstats_groups(const float* x, int ld, long long M, int C, int groups, double* acc, void* stream);
int m2d_bn_apply_groups(const float* x, int ldx, float* y, int ldy, long long M, int C, int groups,
                        const double* acc, const float* gamma, const float* beta,
                        float* running_mean, float* running_var, float momentum, float eps,
                        float* mr, int act, void* stream);
/* stats + apply in ONE launch (grid-wide rendezvous inside the kernel): acc = 2C + 1 ZEROED doubles (sums, sums of
 * squares, rendezvous counter).  What the generator forward uses; colstats / bn_apply remain for callers that
 * accumulate statistics over several calls. */
int m2d_bn_train(const float* x, int ldx, float* y, int ldy, long long M, int C, double* acc,
                 const float* gamma, const float* beta, float* running_mean, float* running_var,
                 float momentum, float eps, float* mr, int act, void* stream);
int m2d_bn_eval(const float* x, int ldx, float* y, int ldy, long long M, int C,
                const float* gamma, const float* beta, const float* running_mean,
                const float* running_var, float eps, int act, void* stream);
int m2d_bn_bwd_reduce(const float* dy, int lddy, const float* y, int ldy, const float* x, int ldx,
                      long long M, int C, const float* mr, int act, double* acc, void* stream);
int m2d_bn_bwd_apply(const float* dy, int lddy, const float* y, int ldy, const float* x, int ldx,
                     float* dx, int lddx, long long M, int C, const float* mr, const float* gamma,
                     int act, const double* acc, float* dgamma, float* dbeta, void* stream);

/* column sums (bias gradients): out[c] = beta*out[c] + scale * sum_m x[m,c] */
int m2d_colsum(const float* x, int ld, long long M, int C, float* out, float scale, float beta,
               double* acc, void* stream);

/* column sums of n matrices in one launch (+ one finalize launch): every bias gradient of a network after a
 * backward sweep.  `table` lives in DEVICE memory; acc = fp64 scratch, acc_off = the entry's offset in it
 * (entries must not overlap; the finalize kernel leaves their slots zero). */
typedef struct {
    const float* x; float* out;
    long long M; long long acc_off;
    int ld, C;
    float scale, beta;
} m2d_colsum_desc;
int m2d_colsum_batch(const m2d_colsum_desc* table, int n, int max_C, double* acc, void* stream);

/* ------------------------------------------------------------------------
 * element-wise / reductions
 * ---------------------------------------------------------------------- */
/* y = a*x (+ b*z if z) */
int m2d_axpby(const float* x, const float* z, float* y, long long n, float a, float b, void* stream);
int m2d_fill(float* y, long long n, float v, void* stream);
/* y[b,:] = s[b]*x[b,:] */
int m2d_scale_rows(const float* x, const float* s, float* y, int nb, long long per, void* stream);
/* losses.py:15-20: xi[b,:] = alpha[b]*real[b,:] + (1-alpha[b])*fake[b,:] */
int m2d_interp(const float* real, const float* fake, const float* alpha, float* xi,
               int nb, long long per, void* stream);
/* the same plus the two copies a critic iteration stacks behind the interpolates (train.py:204-211 evaluates the
 * critic on interpolates, real and fake poses): xi3 = [interpolates (nb); real (nb); fake (nb)] in one pass */
int m2d_interp_stack3(const float* real, const float* fake, const float* alpha, float* xi3,
                      int nb, long long per, void* stream);
/* out[b] += sum x[b,:]^2  (fp64 accumulators) */
int m2d_rows_sumsq(const float* x, int nb, long long per, double* out, void* stream);
/* out[0] += sum x  (fp64) */
int m2d_sum(const float* x, long long n, double* out, void* stream);
/* losses.py:55-60.  ss0/ss1 [B] sums of squares (ss1 NULL when ablated):
 * scal[0] = gp, kappa0[b] = dGP/d||.|| chain factor (2/B)(n-1)/n, same for kappa1 */
int m2d_gp_finalize(const double* ss0, const double* ss1, int B, float* gp, float* kappa0,
                    float* kappa1, float kscale /* kappa *= kscale (the trainer folds gamma in) */, void* stream);
/* losses.py:47-50 (lp=True, the phase2 trainers): gp = mean(max(0, ||g|| - 1)^2), kappa0[b] = (2/B) max(0, n-1)/n */
int m2d_gp_finalize_lp(const double* ss0, int B, float* gp, float* kappa0, void* stream);
/* train.py:226,233 + losses.py:76-82 on channels-last poses [B,T,C]:
 * acc[0] += sum|real-fake|, acc[1] += sum|f[t+1]-f[t]|;
 * dfake = (+= if accumulate) beta*dL1/dfake + eta*dTV/dfake */
int m2d_pose_losses(const float* real, const float* fake, float* dfake, int B, int T, int C,
                    float beta, float eta, int accumulate, double* acc, void* stream);
/* losses.py:85-89 (jerkiness, the smoothness metric of phase3/test.py:85-100) on channels-last poses [B,T,C]:
 * acc[0] += sum over b, t < T-3, c of (x[t+3] - 3 x[t+2] + 3 x[t+1] - x[t])^2;  jerkiness = acc[0] / (B (T-3)) */
int m2d_jerkiness(const float* x, int B, int T, int C, double* acc, void* stream);
/* relu / leaky / tanh derivative applied in place from stored activations */
int m2d_act_bwd(float* d, const float* y, long long n, int mask_mode, void* stream);
/* Input pipeline: SequenceDataset.__getitem__ + collate_fn (utils.py:91-101,128-144) and the H2D
 * copies of phase3/train.py:191-193 on a DEVICE-RESIDENT dataset.  poses / music hold every sequence back to back
 * (float32; offsets in floats); batch entry b becomes frames [start[b], start[b]+T) of sequence seq[b]
 * (real[b, t, :O]) and samples [start[b]*ratio, +A) of its music (audio[b, :A]).  Bit-exact (pure indexing); the
 * caller guarantees start[b] + T <= frames of the sequence and start[b]*ratio + A <= samples of its music. */
int m2d_crop_batch(const float* poses, const long long* pose_off, const float* music, const long long* music_off,
                   const int* seq, const int* start, int B, int T, int O, int ratio, int A, float* real,
                   float* audio, void* stream);
/* out = alpha * a * b * c elementwise on strided row matrices [M][C]: the second-order term of the
 * gradient penalty through a tanh code activation (autograd double backward of default.py:333,303
 * inside losses.py:40-44) */
int m2d_mul3(const float* a, int lda, const float* b, int ldb, const float* c, int ldc, float* out, int ldo,
             long long M, int C, float alpha, void* stream);
/* Label conditioning of the phase2 conditional networks (phase2/archis/conditional.py:19-21 generator,
 * :45-46 critic): nn.Embedding(4,4) lookup + expand over the T rows of a sequence + torch.cat, written directly
 * into the label columns y[(b*T+t)*ldy + e] of the concatenated channels-last operand; labels are int64 like the
 * reference's LongTensor.  *err (device int, may be NULL) is set to 1 when a label is outside [0, n_classes)
 * (the reference raises an index error).  m2d_embed_grad = backward of the lookup:
 * dtable[c,e] = beta*dtable[c,e] + scale * sum over the rows of the sequences labelled c (deterministic). */
int m2d_embed_rows(const float* table, const long long* labels, float* y, int ldy, int B, int T, int E,
                   int n_classes, int* err, void* stream);
int m2d_embed_grad(const float* dy, int ldd, const long long* labels, float* dtable, int B, int T, int E,
                   int n_classes, float scale, float beta, void* stream);
/* MaxPool1d(2,2) and Upsample(x2, linear, align_corners=False) on channels-last rows
 * (default.py:236-237) */
int m2d_maxpool2(const float* x, int ldx, float* y, int ldy, int nb, int Lin, int C, void* stream);
int m2d_maxpool2_bwd(const float* x, int ldx, const float* dy, int lddy, float* dx, int lddx,
                     int nb, int Lin, int C, int accumulate, void* stream);
int m2d_upsample2(const float* x, int ldx, float* y, int ldy, int nb, int Lin, int C, void* stream);
int m2d_upsample2_bwd(const float* dy, int lddy, float* dx, int lddx, int nb, int Lin, int C,
                      int accumulate, void* stream);
/* strided 2-D copy / transpose helpers for the (B,C,L) <-> channels-last boundary */
int m2d_copy2d(const float* x, int ldx, float* y, int ldy, long long M, int C, int accumulate, void* stream);
int m2d_transpose_bcl(const float* x, float* y, int nb, int R, int C, void* stream); /* [b,R,C]->[b,C,R] */
/* n <= M2D_COPY2D_MAX strided copies in one launch, applied in order (`descs` is a HOST array, passed by value to the
 * kernel): the code halves going in and out of the critic's concatenated fusion input (default.py:331-337, torch.cat
 * and its backward), e.g. the audio code of B samples into the three row groups [interpolates; real; fake]. */
#define M2D_COPY2D_MAX 4
typedef struct m2d_copy2d_desc {
    const float* x;     /* source, row stride ldx */
    float* y;           /* destination, row stride ldy */
    long long M;        /* rows */
    int ldx, ldy, C;    /* row strides (floats), columns */
    int accumulate;     /* y += x instead of y = x */
} m2d_copy2d_desc;
int m2d_copy2d_batch(const m2d_copy2d_desc* descs, int n, void* stream);

/* The critic's fusion MLP on n rows (default.py:339-345: fc1 = Linear(F,H) + ReLU, fc2 = Linear(H,1)), one launch:
 *   u[n,H] = relu(W1 x + b1), d[n] = W2 u + b2, and — if dd (upstream of the scores, n floats) is given — the
 *   backward-data dh[n,H] = dd * W2 * [u > 0], dx[n, lddx] = W1^T dh.  w1 (H,F) and w2 (1,H) in the parameter layout.
 *   fp32 CUDA-core arithmetic (the reference's own precision); H <= 256, F <= 8192. */
int m2d_fusion_mlp(const float* x, int ldx, int n, int F, int H, const float* w1, const float* b1,
                   const float* w2, const float* b2, const float* dd, float* u, float* d, float* dh,
                   float* dx, int lddx, void* stream);

/* Scalar losses (train.py:207-214 critic, :226-235 generator).  sums = fp64 {sum D(real),
 * sum D(fake), sum|real-fake|, sum|tv diffs|}.  mode 0: out = {err_fake-err_real+c0*gp, gp,
 * w_dist, err_real, err_fake};  mode 1: out = {err_real-err_fake+c0*l1+c1*tv, l1, tv, err_real, err_fake} */
int m2d_wgan_scalars(const double* sums, const float* gp, int B, long long n_l1, long long n_tv,
                     float c0, float c1, int mode, float* out,
                     const float* d_real, const float* d_fake /* optional: B critic scores each, summed here instead
                                                                 of read from sums[0] / sums[1] */,
                     void* stream);

/* utils.py:329-353 (slice_audio_batch / slice_audio_sequence): out[seq,f,j] =
 * audio[seq, f*stride - pad_left + j], zero outside [0,A).  Pure indexing, bit-exact. */
int m2d_slice_audio(const float* audio, float* out, int nseq, int A, int nwin, int W, int stride,
                    int pad_left, void* stream);

/* torch.optim.Adam (train.py:102-103,216,237): flat fused step over n floats.
 * `step` is a device int32 counter incremented by the call (graph-replay safe);
 * gradients are multiplied by `gscale` first (1/world_size after a sum all-reduce). */
int m2d_adam(float* p, const float* g, float* m, float* v, long long n, int* step,
             float lr, float beta1, float beta2, float eps, float gscale, void* stream);

/* ------------------------------------------------------------------------
 * Optimiser step fused with the weight re-layouts.  ONE launch applies torch.optim.Adam (train.py:102-103,216,237;
 * same arithmetic as m2d_adam) to every parameter a table covers AND rewrites the packed / pre-tiled copies of the
 * convolution weights that m2d_rowconv reads (what m2d_pack_batch produces), reading the weight gradients in the
 * tap-major layout m2d_wgrad(packed = 1) leaves them in — so "gradient unpack -> Adam -> re-layout" is one pass over
 * the parameters instead of three dependent launches.
 *
 * The table (DEVICE memory, built once: all pointers are stable) has one item per tile of a weight tensor
 * (Cout, Cin, k): rows [co0, co0+nco) x channels [ci0, ci0+nci) x taps [t0, t0+nt), nco <= 32, or per plain range of
 * `flat_n` floats (biases, BatchNorm, GRU recurrent weights, ...).  Each thread block stages its tile in shared
 * memory (nco * ((nci * nt) | 1) floats <= `smem_floats`; nt <= 32), so every global access — gradient (tap-major),
 * p / m / v (parameter layout), each packed destination — is made in that array's own contiguous order.
 * pk[i] = the re-layouts of the tile (kinds and destination layouts exactly as in m2d_pack_desc).
 * counters[0] = Adam step count (device int, incremented by the launch: graph-replay safe), counters[1] = scratch (0).
 * ---------------------------------------------------------------------- */
typedef struct {
    float* dst; float* dst_tiled;
    int kind, stride, reserved, pad_;
} m2d_adam_pack_out;
typedef struct {
    float* p; float* m; float* v; const float* g;    /* item base pointers: the LAYER's slices (flat: the range's) */
    long long flat_n;                                 /* > 0: plain range, the fields below are ignored */
    int Cout, Cin, k;
    int g_packed;                                     /* 1: g is tap-major [co, t*Cin + ci]; 0: parameter layout */
    int co0, nco, ci0, nci, t0, nt;
    int n_pack;
    int pad_;                                         /* 1: Cout, Cin, co0, nco, ci0, nci are all multiples of 4 and every
                                                         pointer is 16-byte aligned -> 16-byte accesses throughout */
    m2d_adam_pack_out pk[3];
} m2d_adam_item;
int m2d_adam_pack(const m2d_adam_item* items, int n, int smem_floats, int* counters, float lr, float beta1,
                  float beta2, float eps, float gscale, void* stream);

/* ------------------------------------------------------------------------
 * Gradient all-reduce over NVLink / NVSwitch peer memory (the data-parallel step of SURVEY 8e; the reference has no
 * multi-GPU path).  In-place sum of floats [off, off+n) of a buffer that every GPU of the node has mapped:
 *   bufs[r]        rank r's mapping of the buffer as seen from THIS process (bufs[rank] = the local pointer)
 *   mc             multicast address of the buffer (NVLink SHARP: multimem.ld_reduce / multimem.st), or NULL: the
 *                  reduction is done with peer loads in rank order (deterministic) and peer stores
 *   signal_pads[r] rank r's flag array (unsigned, zero-initialised, >= slot0 + blocks*world entries): block-wise
 *                  release / acquire handshakes before and after the reduction; the flags return to zero
 * ONE kernel of `blocks` thread blocks (two-shot: rank r reduces slice r and broadcasts it); every rank must launch
 * the same sequence of calls.  A peer that does not show up within 2 s sets *status (device int) to 1 instead of
 * hanging the GPU.  Caller contract: producers of the local buffer precede the call on `stream`.
 * ---------------------------------------------------------------------- */
int m2d_nvl_allreduce(float* const* bufs, float* mc, unsigned int* const* signal_pads, int rank, int world,
                      long long off, long long n, int blocks, int slot0, int* status, void* stream);
/* the same over TWO ranges (of possibly different mapped buffers) between one pair of handshakes: the small
 * parameter-layout gradients and the tap-major arena of a network in one launch; the flags of the first buffer */
int m2d_nvl_allreduce2(float* const* bufs, float* mc, long long off, long long n, float* const* bufs2, float* mc2,
                       long long off2, long long n2, unsigned int* const* signal_pads, int rank, int world, int blocks,
                       int slot0, int* status, void* stream);

/* Diagnostics (tools/step_timeline.py): *slot = %globaltimer (ns) when `stream` reaches this point; graph-capturable,
 * so the replayed train step can be cut into phases without a profiler attached. */
int m2d_timestamp(unsigned long long* slot, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* M2D_H */
