/*
 * esr_b200.h -- C ABI of libesr_b200.so, the B200 (sm_100a) implementation of WarranWeng/ESR's
 * per-timestep hot path.  Plain pointers and sizes only; every pointer is a DEVICE pointer unless
 * its name ends in _host.  The caller owns every buffer and the stream; the library allocates no
 * user-visible memory (workspaces are sized by *_workspace_bytes and passed in).  All functions return
 * ESR_OK (0) or a negative error code; esr_last_error() gives the message for the calling thread.
 * Calls are asynchronous on `stream` unless stated otherwise.
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference root).
 */
#ifndef ESR_B200_H
#define ESR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ESR_OK 0
#define ESR_EINVAL (-1)       /* bad argument */
#define ESR_ENEGCOUNT (-2)    /* negative rounded count: the reference raises ValueError (cnt2event.pyx:71) */
#define ESR_ECUDA (-3)        /* CUDA runtime / driver error */
#define ESR_EUNSUPPORTED (-4) /* configuration outside what the sm_100a kernels implement */
#define ESR_EWORKSPACE (-5)   /* workspace too small */

typedef void *esr_stream_t; /* cudaStream_t */

int esr_version(void);
const char *esr_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches claim) */
int64_t esr_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * events -> 2-channel polarity count images
 * Replaces: dataloader/encodings.py:289-304 events_to_channels (+ :243-268 events_to_image), and, with
 * lift_* set, the LR->HR coordinate lift of dataloader/h5dataset.py:508-528 (x / W_lr * W_hr in fp32,
 * two roundings).  F frames in one launch: frame f owns events [frame_off[f], frame_off[f+1]).
 *   xs, ys, ps : fp32 [n_total]        frame_off : int64 [F+1] (device)
 *   out        : fp32 [F, 2, H, W], overwritten (zeroed by the call)
 *   lift_w_lr/lift_w_hr/lift_h_lr/lift_h_hr : 0 = coordinates used as given
 *   writeback  : 1 = reproduce the reference's in-place side effect (out-of-range xs, ys set to 0)
 * Reference quirks kept: out-of-range positive events are dropped, out-of-range NEGATIVE events are
 * counted at neg[0,0]; fractional coordinates truncate toward zero; each event adds ps*ps.
 * writeback: 0 = leave xs / ys alone, 1 = zero them in place for out-of-range events (standalone events_to_channels), 2 = the
 * order of H5Dataset.__getitem__ (h5dataset.py:337-354): in frames of more than 3 events out-of-range events of EITHER polarity add
 * nothing, because create_stack_encoding has sanitised the event tensor before the count encodings see it.
 * --------------------------------------------------------------------------------------------- */
int esr_scatter_cnt(float *xs, float *ys, const float *ps, const int64_t *frame_off, int F, int64_t n_max_frame,
                    int H, int W, int lift_w_lr, int lift_w_hr, int lift_h_lr, int lift_h_hr, int writeback,
                    float *out, esr_stream_t stream);

/* Replaces: dataloader/encodings.py:243-268 events_to_image.  One image, raw weights: out[(long)y,(long)x] += ps;
 * out-of-range events are dropped and, with writeback = 1, xs/ys/ps are zeroed in place like the reference. */
int esr_scatter_image(float *xs, float *ys, float *ps, int64_t n, int H, int W, int writeback, float *out,
                      esr_stream_t stream);

/* Replaces: dataloader/encodings.py:307-331 events_to_mask (index_put_ with accumulate=False: the last event hitting a
 * pixel writes |ps|; out-of-range events are zeroed in place and then write 0 to pixel (0,0)).  last_tmp: int32 [H*W]. */
int esr_scatter_mask(float *xs, float *ys, float *ps, int64_t n, int H, int W, int writeback, int32_t *last_tmp, float *out,
                     esr_stream_t stream);

/* Replaces: the slicing of dataloader/encodings.py:204-240 events_to_stack_no_polarity, i.e. its calls to
 * binary_search_torch_tensor (encodings.py:77-99).  ts: sorted fp32 [n]; bounds: int64 [B,2] = (beg, end) of every time
 * bin, identical to the reference's search (incl. which of several equal timestamps it stops on). */
int esr_time_bin_bounds(const float *ts, int64_t n, int B, int64_t *bounds, esr_stream_t stream);

/* Replaces: dataloader/encodings.py:271-286 events_to_voxel (temporal bilinear voxel grid through events_to_image).
 * out: fp32 [num_bins,H,W], overwritten.  Keeps the reference's side effect (xs, ys zeroed in place for out-of-range
 * events) and its consequence (those events land on pixel (0,0) in bins >= 1). */
int esr_scatter_voxel(float *xs, float *ys, const float *ts, const float *ps, int64_t n, int num_bins, int H, int W,
                      int writeback, float *out, esr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * dense counts / time-bin stacks -> time-sorted event lists
 * Replaces: dataloader/cython_cnt2event/cnt2event.pyx:18-116 (kind 0: vals = [B,2,H,W], P=2, C=1) and
 * dataloader/cython_event_redistribute/event_redistribute.pyx:17-83 / :88-153 (kind 1: vals =
 * [B,P,C,H,W], P=1 for the NoPolarity form).
 *
 * Two phases, because the output length depends on the data:
 *  1. esr_expand_count: rounds (half-to-even) and counts.  stats : int64 [B,4] (device) =
 *     {sum of rounded values, number of events, any-negative flag, max per-slot count}; counts : uint32
 *     [B*P*C*H*W] (device, kept for phase 2).
 *  2. the host reads stats, applies the reference's emptiness rules (a sample whose rounded values sum to
 *     zero yields one zero row; an all-zero call yields [B,1,4]), sizes out = [B, maxlen, 4] (zero-filled)
 *     and calls esr_expand_emit with active_host[b] / start_host[b] (first sorted row of sample b in the
 *     global event order).  mode 0 = linear timestamps (float64 linspace -> fp32), mode 1 = the caller's
 *     random stream rnd[total_events] (float64, numpy MT19937 seed 123, one value per event in emission order).
 * --------------------------------------------------------------------------------------------- */
int esr_expand_count(const float *vals, int B, int P, int C, int H, int W, int kind, int64_t *stats, uint32_t *counts,
                     esr_stream_t stream);
size_t esr_expand_workspace_bytes(int B, int P, int C, int H, int W, int64_t total_events);
/* rank_table (optional, device uint16 [(rank_m+1) * rank_m]): compact sort keys for cnt2event with linear timestamps and
 * max per-pixel count rank_m <= 255 -- rank_table[n*rank_m + j] = index of float32(linspace(0,1,n)[j]) among the sorted
 * distinct timestamps, rank_bits = bits needed; NULL = sort on the raw fp32 timestamp bits (always valid). */
int esr_expand_emit(const float *vals, uint32_t *counts, int B, int P, int C, int H, int W, int kind, int mode,
                    const double *rnd, const uint16_t *rank_table, int rank_m, int rank_bits, const int32_t *active_host,
                    const int64_t *start_host, int64_t total_events, int64_t maxlen, float *out, void *workspace,
                    size_t workspace_bytes, esr_stream_t stream);

/* cnt2event with linear timestamps, whole operator in three launches without a host round trip
 * (dataloader/cython_cnt2event/cnt2event.pyx:18-116, mode 'linear'; csrc/expand_fused.cu).  vals: device fp32 [B,2,H,W].
 * The caller supplies `out` with room for cap_rows rows of 4 floats; the kernels compute maxlen = max over samples of the
 * event count (1 for a sample whose rounded values sum to zero) on the device and write the padded [B, maxlen, 4] result,
 * zero padding included, at the start of `out`.  stats (device int64 [B,4], same meaning as esr_expand_count) is how the caller
 * learns maxlen afterwards and whether the result is valid: it is NOT when an active sample holds a negative count (the
 * reference raises), when the largest count of an active sample exceeds max_count rounded up to a power of two (max_count <= 64:
 * the caller's guess, it sizes the per-key counters in shared memory), when B * maxlen > cap_rows, or when no sample is
 * active -- then nothing was written and the caller applies the reference's rules / uses esr_expand_count + esr_expand_emit.
 * tables: device blob of the 7 key tables for m = 1, 2, 4 .. 64 -- rank uint16 [(m+1)*m] (index of float32(linspace(0,1,n)[j])
 * among the K distinct values for counts <= m) and uniq float [K]; tables_desc_host: int32 [7][3] = {rank byte offset, uniq
 * byte offset, K}.  The host builds them with numpy.linspace, the reference's own arithmetic. */
size_t esr_cnt2event_fused_workspace_bytes(int B, int H, int W);
int esr_cnt2event_fused(const float *vals, int B, int H, int W, const void *tables, const int32_t *tables_desc_host,
                        int max_count, int64_t *stats, float *out, int64_t cap_rows, void *workspace, size_t workspace_bytes,
                        esr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Tensor-core convolution (tcgen05, TMA-tiled implicit GEMM), layer-level entry point.
 * Replaces: every stride-1 nn.Conv2d at feature resolution in models/model.py / models/submodules.py
 * (3x3 pad 1 or 1x1; Cin a multiple of 64, Cout <= 256), including torch.cat inputs (K-split over up to 3
 * sources), the fused bias / residual / activation tail of ConvLayer / ResidualBlock
 * (models/submodules.py:159-200, 347-409) and the ConvGRU gating (models/submodules.py:496-514).
 *
 * Activation tensors are "split bf16" NHWC: [2 planes][n_img][H][W][C] bf16, value = hi + lo
 * (esr_split_from_nchw / esr_split_to_nchw convert from / to the reference's fp32 NCHW layout).
 * Weights are packed once by esr_pack_conv_weight (fp32 [Cout,Cin,k,k] -> split bf16 K-blocks).
 * --------------------------------------------------------------------------------------------- */
typedef struct esr_conv_desc {
    const void *src[3];        /* split tensors */
    int src_C[3];              /* channels of each source (multiple of 64) */
    int src_n_img[3];          /* images in each source tensor */
    const int32_t *src_img[3]; /* optional: output image -> source image index (device), NULL = identity */
    int n_src;
    int H, W;                  /* spatial size (input == output) */
    int n_img;                 /* output images */
    int ntaps;                 /* 9 = 3x3 pad 1, 1 = 1x1 */
    int cout;
    const void *wpacked;       /* from esr_pack_conv_weight */
    const float *bias;         /* fp32 [ceil16(cout)], zero padded */
    int act;                   /* 0 none, 1 relu, 2 sigmoid, 3 tanh; applied to channels >= act_from */
    int act_from;
    int res_mode;              /* 0 none, 1 add before activation, 2 add after activation */
    int epi_mode;              /* 0 standard, 1 ConvGRU update|reset gates, 2 ConvGRU candidate + blend */
    const void *res; int res_C; int res_n_img; const int32_t *res_img;
    void *out; int out_C; int out_n_img; int out_coff;   /* split output (NULL = none), channel offset */
    float *out_f32; int out_f32_C;                        /* fp32 NHWC output (NULL = none) */
    const void *h_prev; int h_n_img; float *z_buf;        /* ConvGRU: previous state (split, 64 ch), z gate fp32 */
} esr_conv_desc;

int esr_conv_tc(const esr_conv_desc *desc, esr_stream_t stream);
size_t esr_conv_weight_bytes(int cout, int cin, int ksz);
/* w1 != NULL: two [cout_each,cin,k,k] tensors concatenated along Cout (GRU update|reset) */
int esr_pack_conv_weight(const float *w0, const float *w1, int cout_each, int cin, int ksz, void *dst,
                         esr_stream_t stream);
int esr_split_from_nchw(const float *src, int n_img, int C, int H, int W, void *dst, esr_stream_t stream);
int esr_split_to_nchw(const void *src, int n_img, int C, int H, int W, float *dst, esr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * The network: DeepRecurrNet.forward with carried ConvGRU states.
 * Replaces: models/model.py:294-344 (DeepRecurrNet.forward / reset_states), and underneath it
 * models/model.py:20-291, models/submodules.py (ConvLayer, UpsampleConvLayer, ResidualBlock, RecurrentConvLayer,
 * ConvGRU, MLP), models/model_util.py:133-164 (CropSize) and the `_ext.dcn_v2_forward` operator
 * (models/DCNv2/src/dcn_v2.h:9-27, src/cuda/dcn_v2_cuda.cu:20-95) for the shipped configuration
 * (inch=2, basech=8, num_frame=3, norm=None, relu, all sub-blocks enabled; config/train_ours_enfssyn.yml:21-26).
 *
 *  params    : 68 device pointers to fp32 tensors in the reference's state_dict order
 *              (head.conv2d.weight, head.conv2d.bias, feat_extract.convblock.0.conv2d.weight, ... tail.conv2d.bias)
 *  blob      : esr_net_param_bytes() bytes; repack whenever the parameters change
 *  workspace : esr_net_workspace_bytes(B,N,L,H,W) bytes, owned by the caller, must outlive the net; holds all
 *              intermediates AND the recurrent states (which persist across esr_net_forward calls)
 *  L         : frames per sequence handled by one call.  L == N (=3) is the reference's forward: one window,
 *              input fp32 [B,N,2,H,W], output fp32 [B,2,H,W].  L > N is the sequence form used by the pipeline:
 *              input fp32 [B,L,2,H,W], output fp32 [(L-N+1)*B,2,H,W] (window-major: w*B+b) = the L-N+1 sliding-window
 *              forwards the reference would run one after another with the state carried (train_ours_cnt_seq.py:217-231,
 *              dataloader/h5dataloader.py:229-231).  Per-frame work (encoder, attention maps) is done once per frame
 *              and all state-independent layers once for all windows; only the ConvGRU chain is serial.  Results are
 *              identical to L-N+1 single-window calls.
 *  in_img    : optional device int32 [B*L]: frame (b,l) is read from input image in_img[b*L+l] (frame banks)
 * H, W need not be multiples of 8: the CropSize pad / crop is folded into the first and last kernels.
 * --------------------------------------------------------------------------------------------- */
typedef void *esr_net_t;
size_t esr_net_param_bytes(void);
int esr_net_pack_params(const float *const *params_host_array_of_device_ptrs, void *blob, esr_stream_t stream);
size_t esr_net_workspace_bytes(int B, int N, int L, int H, int W);
int esr_net_create(esr_net_t *net, int B, int N, int L, int H, int W, void *blob, void *workspace, size_t workspace_bytes,
                   esr_stream_t stream);
int esr_net_destroy(esr_net_t net);
int esr_net_reset_states(esr_net_t net, esr_stream_t stream);
int esr_net_forward(esr_net_t net, const float *input, const int32_t *in_img, float *output, esr_stream_t stream);
/* Same as esr_net_forward, but brackets every kernel launch with CUDA events on `stream`, synchronises, and reports
 * per-launch {class (0 tensor-core conv, 1 CUDA-core conv, 2 element-wise/sampling, 3 cooperative ConvGRU chain),
 * milliseconds, algorithmic FLOPs, algorithmic bytes (every input / output element once at 4 bytes), layer name (32 chars each)}
 * into host arrays (measurement aid for bench.py's roofline objects; not on the production path).  bytes_host / names_host
 * may be null. */
int esr_net_forward_profiled(esr_net_t net, const float *input, const int32_t *in_img, float *output, int max_entries,
                             int *n_entries_host, int *cls_host, float *ms_host, double *flops_host, double *bytes_host,
                             char *names_host, esr_stream_t stream);
/* states: fp32 [2,B,64,H/8,W/8] (forward-direction state, reverse-direction state), the reference's self.states */
int esr_net_get_states(esr_net_t net, float *states, esr_stream_t stream);
int esr_net_set_states(esr_net_t net, const float *states, esr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * The `_ext.dcn_v2_forward` operator (modulated deformable 3x3 convolution), reference layouts in and out.
 * Replaces: models/DCNv2/src/dcn_v2.h:9-27 dcn_v2_forward -> src/cuda/dcn_v2_cuda.cu:20-95 (+ the im2col kernel
 * src/cuda/dcn_v2_im2col_cuda.cu:125-195), as bound by src/vision.cpp:4-8 and called from models/DCNv2/dcn_v2.py:27.
 * input [B,C,H,W], weight [Co,C,3,3], bias [Co], offset [B,dg*18,H,W], mask [B,dg*9,H,W], output [B,Co,H,W]; fp32.
 * Implemented configuration: the one ESR instantiates (C=Co=64, 3x3, stride 1, pad 1, dilation 1, dg=8);
 * anything else returns ESR_EUNSUPPORTED.
 * --------------------------------------------------------------------------------------------- */
size_t esr_dcn_v2_workspace_bytes(int B, int H, int W);   /* the configuration ESR uses (64 -> 64, 3x3, s1 p1 d1, 8 groups) */
/* workspace for ANY configuration the reference operator accepts (backward = 1: for esr_dcn_v2_backward).  The configuration
 * of models/model.py:173 runs on the tcgen05 path; every other one (the reference's own tests use 2 -> 2 channels and
 * deformable_groups 1 / 2, models/DCNv2/testcuda.py:14-17,169-180) on fp32 CUDA-core kernels (csrc/dcn_generic.cu). */
size_t esr_dcn_v2_workspace_bytes_ex(int B, int C, int H, int W, int Co, int kernel, int stride, int pad, int dilation,
                                     int deformable_group, int backward);
/* Replaces: models/DCNv2/src/dcn_v2.h:29-50 dcn_v2_backward -> src/cuda/dcn_v2_cuda.cu:97-216 (+ the col2im / coord kernels
 * src/cuda/dcn_v2_im2col_cuda.cu:197-327), called from models/DCNv2/dcn_v2.py:50.  Same five gradients, reference layouts:
 * grad_input [B,C,H,W], grad_offset [B,dg*18,H,W], grad_mask [B,dg*9,H,W], grad_weight [Co,C,3,3], grad_bias [Co].
 * grad_input uses fp32 atomics like the reference (summation order is not deterministic). */
size_t esr_dcn_v2_backward_workspace_bytes(int B, int H, int W);
int esr_dcn_v2_backward(const float *input, const float *weight, const float *bias, const float *offset, const float *mask,
                        const float *grad_output, int B, int C, int H, int W, int Co, int kernel, int stride, int pad,
                        int dilation, int deformable_group, float *grad_input, float *grad_offset, float *grad_mask,
                        float *grad_weight, float *grad_bias, void *workspace, size_t workspace_bytes, esr_stream_t stream);
int esr_dcn_v2_forward(const float *input, const float *weight, const float *bias, const float *offset, const float *mask,
                       int B, int C, int H, int W, int Co, int kernel, int stride, int pad, int dilation,
                       int deformable_group, float *output, void *workspace, size_t workspace_bytes, esr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Training step operators.
 * Replaces, for one ConvLayer (models/submodules.py:159-200: Conv2d(bias) -> activation), ATen's conv2d forward and
 * backward as autograd calls them inside train_ours_cnt_seq.py:217-231 (forward) and :232 (loss.backward()); and
 * nn.MSELoss (:774) and torch.optim.Adam(lr, weight_decay, amsgrad) (:781, config/train_ours_enfssyn.yml optimizer).
 * Tensors are fp32 NCHW (x [B,Cin,H,W], w [Cout,Cin,k,k], y/dy [B,Cout,Ho,Wo]); k = 3 (pad 1) or 1 (pad 0); stride 1|2;
 * act: 0 none, 1 relu, 2 sigmoid, 3 tanh (fused into the forward; backward multiplies dy by act'(y)).
 * backward: dx may be NULL (first layer); dw [Cout,Cin,k,k] and db [Cout] are overwritten (not accumulated); dw == db ==
 * NULL computes dx only (a caller that batches the weight gradient of a weight-shared layer, e.g. the ConvGRU steps).
 * workspace: esr_conv2d_workspace_bytes() bytes of device memory owned by the caller.
 * --------------------------------------------------------------------------------------------- */
size_t esr_conv2d_workspace_bytes(int B, int Cin, int H, int W, int Cout, int ksz, int stride);
/* x_split (optional): layers whose forward, dx and dw all run on the tensor cores convert x to the split-bf16 NHWC operand
 * format once; esr_conv2d_split_bytes() > 0 says so and gives the size of the buffer the forward fills (x_split_out) and the
 * backward reads (x_split) instead of converting x again -- the backward then does not need x itself (x may be NULL). */
size_t esr_conv2d_split_bytes(int B, int Cin, int H, int W, int Cout, int ksz, int stride);
int esr_conv2d_forward(const float *x, const float *w, const float *bias, int B, int Cin, int H, int W, int Cout, int ksz,
                       int stride, int act, float *y, void *x_split_out, void *workspace, size_t workspace_bytes,
                       esr_stream_t stream);
int esr_conv2d_backward(const float *x, const void *x_split, const float *w, const float *y, const float *dy, int B, int Cin,
                        int H, int W, int Cout, int ksz, int stride, int act, float *dx, float *dw, float *db, void *workspace,
                        size_t workspace_bytes, esr_stream_t stream);
/* Bilinear x2 upsampling of `planes` = B*C fp32 planes [H,W] -> [2H,2W] and its backward (dy [2H,2W] -> dx [H,W]):
 * F.interpolate(scale_factor=2, mode='bilinear', align_corners=False) of UpsampleConvLayer (models/submodules.py:290). */
int esr_upsample2x_forward(const float *x, int planes, int H, int W, float *y, esr_stream_t stream);
int esr_upsample2x_backward(const float *dy, int planes, int H, int W, float *dx, esr_stream_t stream);
/* Resize of `planes` fp32 planes [Hin,Win] -> [Hout,Wout] as torch.nn.functional.interpolate(size=..., align_corners=False)
 * does on the CPU: mode 1 = 'bicubic' (Keys, A = -0.75, clamped taps), mode 0 = legacy 'nearest'.  Replaces the per-frame
 * calls of the dataset's tensor factory (dataloader/h5dataset.py:341-344: inp_bicubic_cnt / _stack, inp_near_cnt / _stack)
 * and the bicubic baseline of infer_ours_cnt.py:76-78. */
int esr_resize_planes(const float *x, int planes, int Hin, int Win, int Hout, int Wout, int mode, float *out,
                      esr_stream_t stream);
/* ConvGRU gate arithmetic (models/submodules.py:507-512), fp32, B images of chw = C*H*W elements; zr [B, 2C, H, W] holds
 * the update gate z in channels [0,C) and the reset gate r in [C,2C).  hr = h*r;  blend = h*(1-z) + o*z.  The backward
 * entry points write full-size dzr (the half they do not touch is zero-filled). */
int esr_gru_hr(const float *h, const float *zr, int B, int chw, float *out, esr_stream_t stream);
int esr_gru_hr_backward(const float *h, const float *zr, const float *grad, int B, int chw, float *dh, float *dzr,
                        esr_stream_t stream);
int esr_gru_blend(const float *h, const float *zr, const float *o, int B, int chw, float *out, esr_stream_t stream);
int esr_gru_blend_backward(const float *h, const float *zr, const float *o, const float *grad, int B, int chw, float *dh,
                           float *dzr, float *d_o, esr_stream_t stream);
/* loss[0] = mean((pred - target)^2); grad (optional) = grad_scale * 2 (pred - target) / n */
int esr_mse_loss(const float *pred, const float *target, size_t n, float *loss, float *grad, float grad_scale,
                 esr_stream_t stream);
/* One torch.optim.Adam step over a flat fp32 parameter buffer; max_exp_avg_sq != NULL = amsgrad.  step_counter is a
 * DEVICE int32 holding the number of steps taken so far (0 before the first); the call increments it on the stream and
 * uses the new value for the bias corrections, so the call can sit inside a replayed CUDA graph. */
int esr_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, float *max_exp_avg_sq, size_t n,
                  int32_t *step_counter, float lr, float beta1, float beta2, float eps, float weight_decay,
                  esr_stream_t stream);
/* Same step with the hyper-parameters {lr, beta1, beta2, eps, weight_decay} read from DEVICE memory when the kernel runs: a
 * CUDA graph that contains the call follows a learning-rate schedule (train_ours_cnt_seq.py:784) by rewriting 20 bytes. */
int esr_adam_step_dev(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, float *max_exp_avg_sq, size_t n,
                      int32_t *step_counter, const float *hyper, esr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Evaluation metrics on the GPU.  Replaces the per-channel CPU calls of infer_ours_cnt.py:81-100:
 * nn.L1Loss / nn.MSELoss, loss/restore.py:42-61 ssim_loss (skimage structural_similarity, 7x7 uniform window, sample
 * covariance, K1 0.01, K2 0.03) and :64-90 psnr_loss (skimage peak_signal_noise_ratio).
 * pred, tgt: fp32 [n_planes, H, W] (plane = sample x channel).  stats: fp64 [n_planes][6] =
 *   {sum |pred - tgt|, sum (pred - tgt)^2, max tgt, min tgt, sum of the SSIM map over the valid region, its pixel count};
 * the host side (esr_b200/metrics.py) turns them into the reference's scalars. */
size_t esr_metrics_workspace_bytes(int n_planes, int H, int W, int win);
int esr_metrics_planes(const float *pred, const float *tgt, int n_planes, int H, int W, int win, double data_range, double *stats,
                       void *workspace, size_t workspace_bytes, esr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Columnar event reader, device side.  Replaces, for whole batches of frames:
 *   BaseDataset.binary_search_h5_dset (dataloader/base_dataset.py:78-91; same algorithm as
 *   dataloader/binary_search/binary_search.pyx:17-38) used by H5Dataset.find_ts_index / get_gt_event_indices_num
 *   (dataloader/h5dataset.py:264-270, 451-475): out[i] = the index that bisection returns for queries[i] (exact hit -> the probed
 *   index, else the left insertion point); ts sorted float64 [n], device or host-mapped memory;
 *   H5Dataset.get_events / get_gt_events + BaseDataset.event_formatting (h5dataset.py:492-506, base_dataset.py:26-33):
 *   frame f = rows [start[f], start[f] + off[f+1] - off[f]) of the int16 x / y and float64 t / p columns -> fp32 SoA at
 *   out_*[off[f] ...]; out_ts (optional) = per-frame normalised time (ts - ts[0]) / (ts[-1] - ts[0] + 1e-6) in fp32. */
int esr_ts_search(const double *ts, int64_t n, const double *queries, int64_t nq, int64_t *out, esr_stream_t stream);
int esr_gather_events(const int16_t *xs, const int16_t *ys, const double *ts, const double *ps, const int64_t *start,
                      const int64_t *off, int n_frames, int64_t max_len, float *out_xs, float *out_ys, float *out_ts,
                      float *out_ps, esr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ESR_B200_H */
