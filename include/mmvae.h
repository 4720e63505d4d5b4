/* mmvae.h -- C ABI of libmmvae_b200.so: the B200-native VAE training step.
 *
 * This is the drop-in boundary for ONE hot path of praateekmahajan/moving-mnist-vae:
 * forward + loss + backward of the `VAE` module (reference model.py:258-406) as driven
 * by the train-loop body main.py:389-390,398.  Everything behind these entry points is
 * hand-written CUDA for sm_100a.  There is no CPU fallback: every compute entry point
 * returns MMVAE_ERR_NO_DEVICE / MMVAE_ERR_ARCH instead of computing on the host.
 *
 * Conventions
 *   - plain C, no torch / CUDA types in signatures (`stream` is a cudaStream_t passed
 *     as void*; device pointers are plain pointers);
 *   - the library never allocates, frees or retains device memory: inputs, parameter /
 *     gradient / buffer arenas, outputs and the workspace are owned by the caller (the
 *     Python shim uses torch's caching allocator);
 *   - all work is enqueued on `stream`; no host synchronisation, no default-stream use,
 *     so a sequence of calls is CUDA-graph capturable;
 *   - return 0 on success, a negative MMVAE_ERR_* otherwise; the message is available
 *     through mmvae_last_error() (thread-local);
 *   - public tensors use the reference's layout: NCHW fp32.  Internal activations are
 *     NHWC in the storage type of `precision`.
 *
 * Parameter arena: all 93 (base model) parameter tensors of the reference's
 * `named_parameters()` order, back to back, fp32, no padding (2,112,819 floats for the
 * base model).  The gradient arena has the same layout.  The BatchNorm buffer arena is
 * [running_mean(C), running_var(C)] per BatchNorm2d in state_dict order; `bn_counters`
 * holds the matching `num_batches_tracked` values (int64).  mmvae_param_entry /
 * mmvae_bn_entry enumerate both so the host never hard-codes offsets.
 */
#ifndef MMVAE_H_
#define MMVAE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMVAE_ABI_VERSION 4

enum {
  MMVAE_OK = 0,
  MMVAE_ERR_BAD_DESC = -1,    /* unsupported / inconsistent mmvae_desc                  */
  MMVAE_ERR_BAD_ARG = -2,     /* NULL / misaligned pointer, workspace too small         */
  MMVAE_ERR_NO_DEVICE = -3,   /* no CUDA device visible                                 */
  MMVAE_ERR_ARCH = -4,        /* device is not sm_100 (B200); nothing else is compiled  */
  MMVAE_ERR_CUDA = -5         /* a CUDA runtime call / launch failed                    */
};

enum { MMVAE_PREC_FP32 = 0,   /* fp32 storage, fp32 SIMT arithmetic: validation mode    */
       MMVAE_PREC_BF16 = 1 }; /* bf16 storage, tcgen05 bf16 MMA with fp32 accumulation  */

enum { MMVAE_LOSS_GAUSSIAN = 0,   /* -nll * Normal(recon, sigma).log_prob(target).sum()  model.py:403     */
       MMVAE_LOSS_CATEGORICAL = 1 /*  nll * cross_entropy(recon, target, w).sum()        model.py:400-401 */ };

/* Mirrors the constructor arguments of the reference VAE (model.py:259-262) that shape the
 * hot path, plus the local batch size and the arithmetic mode. */
typedef struct mmvae_desc {
  int32_t struct_size;      /* = sizeof(mmvae_desc), checked                                   */
  int32_t batch;            /* N: frames in this call (per GPU)                                */
  int32_t in_channels;      /* model.py:259 in_channels                                        */
  int32_t out_channels;     /* model.py:259 decoder_out_channels                               */
  int32_t z_dim;            /* model.py:260 z_dimension                                        */
  int32_t image_size;       /* model.py:262 input_image_size (>32: five up-blocks, else four)  */
  int32_t width;            /* channel multiplier of the literals at model.py:92-101,157-170   */
  int32_t require_rsample;  /* model.py:262; 0 => encoding = mu, no logvar head                 */
  int32_t precision;        /* MMVAE_PREC_*                                                    */
  int32_t training;         /* 1: batch-statistics BatchNorm + running-stat update; 0: eval    */
  int32_t flags;            /* MMVAE_FLAG_* (0 = default kernel selection)                     */
  int32_t arch;             /* MMVAE_ARCH_* (0 = model.py VAE)                                 */
  int32_t reserved[4];
} mmvae_desc;

/* Network family.  MMVAE_ARCH_NOTEBOOK is the BatchNorm-free variant of vae-kl.ipynb:119-166 (BASELINE configs[4]):
 * encoder conv(1->C,k5,s2,p2) ReLU, conv(C->C,k5,s2,p1) ReLU, conv(k3,s2,p1) ReLU x2, heads conv(C->z,k3,s2,p1) x2;
 * decoder up2 conv3x3(z->C) ELU, up4 conv ELU, up2 conv ELU, up2 conv3x3(C->out_channels) = logits; every conv has a
 * bias.  C = 32*width, out_channels = classes (256 in the notebook, a multiple of 8, <= 256), image_size 64 or 128,
 * in_channels 1; require_rsample / training are ignored (no BatchNorm, always samples).  mu / logvar / encoding are
 * [N, z, h, h] (h = image_size/32).  Parameter arena order: encoder.conv1.weight, .bias, ... decoder.conv4.bias
 * (list(encoder.parameters()) + list(decoder.parameters()), vae-kl.ipynb cell 7). */
enum { MMVAE_ARCH_RESNET = 0, MMVAE_ARCH_NOTEBOOK = 1 };

/* Kernel selection for validation: MMVAE_PREC_BF16 normally runs every layer shape the tcgen05
 * kernels cover on the tensor cores; this flag forces the fp32-FMA SIMT kernels (same bf16 storage)
 * so the two can be compared tensor by tensor. */
enum { MMVAE_FLAG_FORCE_SIMT = 1,
       /* MMVAE_ARCH_NOTEBOOK, bf16: mmvae_forward (called with recon == NULL) stops before decoder.conv4 and
        * mmvae_nb_loss_backward runs that conv fused with the softmax cross-entropy, writing d logits directly: the
        * 256-channel logits never reach HBM.  Both calls must carry the flag. */
       MMVAE_FLAG_DEFER_LOGITS = 2 };

typedef struct mmvae_layout_info {
  int64_t n_params;         /* floats in the parameter / gradient arena         */
  int64_t n_bn_buffers;     /* floats in the BatchNorm buffer arena             */
  int32_t n_param_tensors;  /* 93 for the base model                            */
  int32_t n_bn;             /* 30 for the base model (29 at image_size <= 32)   */
  int64_t workspace_bytes;  /* activations + saved statistics + backward scratch */
  int32_t decoder_size;     /* spatial size the decoder emits (64 or 32) before the crop */
  int32_t crop;             /* `adjust` of model.py:307-310 (>= 0 pixels per side)       */
  int64_t train_flops;      /* 2*MAC over every conv, fwd + dgrad + wgrad, for `batch` frames */
} mmvae_layout_info;

typedef struct mmvae_loss_args {
  int32_t struct_size;      /* = sizeof(mmvae_loss_args)                                        */
  int32_t kind;             /* MMVAE_LOSS_*                                                     */
  float nll;                /* model.py:285 self.nll                                            */
  float kl;                 /* model.py:285 self.kl -- a per-call scalar so it can be annealed  */
  float sigma;              /* model.py:286 sigma_decoder (Gaussian branch)                     */
  int32_t batch;            /* N = target.shape[0]   (model.py:405)                             */
  int32_t channels;         /* C of recon                                                       */
  int32_t height, width;    /* spatial size of recon / target (after the crop)                  */
  int32_t z_dim;            /* elements of mu / logvar per frame                                */
  int32_t reserved;
  const float* kl_dev;      /* NULL, or a DEVICE scalar that overrides `kl`: the KL weight can then be
                             * annealed between replays of a captured CUDA graph                 */
} mmvae_loss_args;

/* ---- introspection: callable without a GPU --------------------------------------------- */
int mmvae_abi_version(void);
const char* mmvae_last_error(void);
int mmvae_layout(const mmvae_desc* d, mmvae_layout_info* out);
/* i-th parameter tensor (reference named_parameters() order): name, arena offset, shape. */
int mmvae_param_entry(const mmvae_desc* d, int32_t i, char* name, size_t name_cap,
                      int64_t* offset, int32_t* ndim, int32_t shape[4]);
/* i-th BatchNorm2d: state_dict prefix, channels, offset of running_mean in the buffer arena
 * (running_var follows at +channels). */
int mmvae_bn_entry(const mmvae_desc* d, int32_t i, char* prefix, size_t prefix_cap,
                   int32_t* channels, int64_t* buffer_offset);
/* Named tensor inside the workspace (debug / parity tests): "encoder.layer1.0.conv1" is the raw
 * conv output y (NHWC, storage type), "encoder.layer1.0" the block output; "<name>.grad" / "<name>.grad2" the gradient
 * buffer(s) of that tensor after mmvae_backward.  dims = {N,H,W,C}.  MMVAE_ERR_BAD_ARG for a tensor the selected kernels
 * never materialise (bf16 tensor-core path: "encoder.conv1.grad" -- the stem weight gradient forms that dY in its loader). */
int mmvae_workspace_tensor(const mmvae_desc* d, const char* name, int64_t* byte_offset, int32_t dims[4]);

/* ---- the hot path ---------------------------------------------------------------------------
 * VAE.forward (model.py:316-342, pixelcnn None): encoder -> rsample -> decoder.
 *   x        [N, in_channels, S, S] fp32 NCHW (device)
 *   params   parameter arena (fp32, device)
 *   bn_buffers / bn_counters   updated in place when d->training (momentum 0.1, unbiased var)
 *   eps      [N, z] fp32 standard-normal draw of rsample (model.py:149-150); NULL => generated on
 *            device with Philox4x32-10 keyed by (seed, offset) and written to eps_out
 *   eps_out  [N, z] fp32 or NULL
 *   rng_state  NULL, or a device pointer to {seed, offset} (uint64 x 2) that overrides the by-value pair:
 *            lets a captured CUDA graph draw fresh noise on every replay.  The call itself advances
 *            rng_state[1] by ceil(N*z/4) after the draw (model.py VAE; the notebook variant leaves it to
 *            the host)
 *   mu, logvar, encoding   [N, z] fp32 outputs (logvar may be NULL iff !require_rsample)
 *   recon    [N, out_channels, D, D] fp32 NCHW, D = decoder_size (crop is a host-side view)
 * The workspace keeps what mmvae_backward needs. */
int mmvae_forward(const mmvae_desc* d, const float* x, const float* params, float* bn_buffers,
                  int64_t* bn_counters, const float* eps, uint64_t seed, uint64_t offset, float* eps_out,
                  uint64_t* rng_state, void* workspace, size_t workspace_bytes,
                  float* mu, float* logvar, float* encoding, float* recon, void* stream);

/* MMVAE_ARCH_NOTEBOOK: mmvae_forward takes the same arguments (bn_buffers / bn_counters unused, may be NULL); `recon`
 * (the logits, [N, classes, S, S] fp32 NCHW) may be NULL: at training batch sizes the logits stay in the workspace in the
 * storage type (4.3 GB at N=512, 128x128) and only mmvae_nb_loss_backward reads them.
 *
 * mmvae_nb_loss_backward = the rest of the notebook's loop body (vae-kl.ipynb:225-231) in one call:
 *   pxz = sum CE(logits, target) / N,  kl = sum KL(N(mu, exp(logvar/2)) || N(0,1)) / N,  loss = pxz + kl_weight * kl
 * and the gradient of `loss` with respect to every parameter (arena `grads`, overwritten).  target: int64 [N, S, S]
 * class indices; out: 3 floats on the device {loss, pxz, kl}.  Must follow a mmvae_forward with the same desc /
 * workspace / x / params.  kl_weight = 1 is the notebook; any other value is the annealed configuration. */
int mmvae_nb_loss_backward(const mmvae_desc* d, const float* x, const int64_t* target, const float* params,
                           void* workspace, size_t workspace_bytes, float kl_weight, float* out, float* grads,
                           void* stream);

/* Decoder only (get_z_image / get_reconstruction, model.py:344-362): encoding [N, z] -> recon. */
int mmvae_decode(const mmvae_desc* d, const float* encoding, const float* params, float* bn_buffers,
                 int64_t* bn_counters, void* workspace, size_t workspace_bytes, float* recon, void* stream);

/* VAE.loss forward (model.py:385-406; the MMD diagnostic is mmvae_mmd): out[0] = (pxz + kl*KL)/N,
 * out[1] = pxz/N, out[2] = KL/N, all fp32 on the device.  ONE launch: the reconstruction term and the
 * KL terms are reduced by the same grid and the CTA that finishes last combines the per-CTA partials in
 * fp64 in a fixed order.  target: fp32 [N,C,H,W] (Gaussian) or int64 [N,H,W] (categorical; a class
 * outside [0,C) yields a NaN loss, never an out-of-bounds read); ce_weight [C] fp32 or NULL.
 * mu/logvar may be NULL (KL = 0).
 * scratch: >= mmvae_loss_scratch_bytes() bytes of device memory, ZERO before its first use (every call
 * leaves it zero again). */
size_t mmvae_loss_scratch_bytes(void);
int mmvae_loss_forward(const mmvae_loss_args* a, const float* recon, const void* target,
                       const float* ce_weight, const float* mu, const float* logvar,
                       float* out, void* scratch, void* stream);
/* Gradient of out[0] scaled by the device scalar *grad_out: d_recon [N,C,H,W], d_mu, d_logvar [N,z]
 * (ONE launch for all three). */
int mmvae_loss_backward(const mmvae_loss_args* a, const float* recon, const void* target,
                        const float* ce_weight, const float* mu, const float* logvar,
                        const float* grad_out, float* d_recon, float* d_mu, float* d_logvar, void* stream);

/* Backward of mmvae_forward: given dL/d(mu, logvar, encoding, recon) (any may be NULL = zero;
 * d_recon is [N, out_channels, D, D] uncropped) fill the gradient arena `grads` (overwritten, fp32,
 * same layout as `params`).  Must follow a mmvae_forward with the same desc / workspace / x / params.
 * `phases` selects which part of the backward sweep to enqueue, so that a data-parallel host can record
 * an event after each part and start that part's gradient all-reduce while the rest still runs; the
 * parts must be issued in order DECODER, ENC_DEEP, ENC_SHALLOW.  Each part owns a contiguous range of
 * the gradient arena (mmvae_backward_range). */
enum { MMVAE_BWD_DECODER = 1,      /* tail conv .. decoder stem: every decoder.* gradient          */
       MMVAE_BWD_ENC_DEEP = 2,     /* heads, encoder.layer4, encoder.layer3                        */
       MMVAE_BWD_ENC_SHALLOW = 4,  /* encoder.layer2, layer1, stem                                 */
       MMVAE_BWD_ALL = 7,
       /* OR-ed into a single DECODER / ENC_DEEP phase: its weight gradients (enqueued on the library's auxiliary stream)
        * are NOT joined into `stream` when the call returns, so the rest of the sweep does not wait for them; whoever
        * consumes that phase's gradients first makes ITS stream wait with mmvae_aux_fence().  The ENC_SHALLOW phase
        * always joins (everything the sweep forked is back on `stream` when it returns). */
       MMVAE_BWD_DEFER_JOIN = 8 };
int mmvae_backward(const mmvae_desc* d, const float* x, const float* params,
                   void* workspace, size_t workspace_bytes,
                   const float* d_mu, const float* d_logvar, const float* d_encoding, const float* d_recon,
                   float* grads, int32_t phases, void* stream);
/* Make `stream` wait for everything this library has enqueued so far on its auxiliary streams of the current device
 * (an event hand-over; nothing blocks on the host).  With MMVAE_BWD_DEFER_JOIN this is how a communication stream gets a
 * phase's weight gradients without stalling the compute stream. */
int mmvae_aux_fence(void* stream);
/* [begin, end) float offsets of the gradient-arena range written by one phase. */
int mmvae_backward_range(const mmvae_desc* d, int32_t phase, int64_t* begin, int64_t* end);

/* Input side of the step (main.py:381-388): k-means label map (uint8, N*H*W) -> normalised fp32 network
 * input x = (label - data_mean) / data_std, and optionally the int64 cross-entropy target. */
int mmvae_prepare_input(const uint8_t* labels, int64_t n, float data_mean, float data_std, float* x, int64_t* target,
                        void* stream);

/* Kernels this library has launched in this process so far (monotonic). */
int64_t mmvae_launch_count(void);

/* Counter-based standard normals, the generator mmvae_forward uses when eps == NULL:
 * element i = Box-Muller of Philox4x32-10(key = seed ^ stream_id, counter = offset + i/4)[i%4 pair].
 * rng_state: NULL, or a device {seed, offset} pair overriding the by-value one (as in mmvae_forward);
 * stream_id = 0 is the rsample stream, any other value an independent stream of the same generator
 * (VAE.loss draws the MMD diagnostic's true_samples, model.py:395, from stream 1). */
int mmvae_philox_normal(uint64_t seed, uint64_t offset, const uint64_t* rng_state, uint64_t stream_id,
                        int64_t n, float* out, void* stream);

/* The MMD diagnostic VAE.loss returns as its 4th value (model.py:367-383,394-396,406):
 *   k(a,b) = exp(-mean_d((a_d-b_d)^2)/dim),  mmd = sum k(x,x) + sum k(y,y) - 2 sum k(x,y),  out[0] = mmd / N
 * with x = true_samples [N,z] and y = encoding [N,z] (fp32, device).  One launch, fp64 combine in a fixed
 * order.  scratch: >= mmvae_mmd_scratch_bytes(N) bytes, ZERO before its first use. */
size_t mmvae_mmd_scratch_bytes(int32_t n);
int mmvae_mmd(const float* true_samples, const float* encoding, int32_t n, int32_t z_dim, float* out,
              void* scratch, void* stream);

/* Fused multi-tensor Adam over the flat arenas (optim.Adam defaults of main.py:468, optimizer.step()
 * main.py:399): p -= lr * m_hat / (sqrt(v_hat) + eps), grads pre-scaled by grad_scale
 * (1/world_size when the gradients were summed, not averaged, over ranks).  `step` is the 1-based step
 * count; step_dev: NULL, or a device int64 that overrides it (bias corrections computed on the device, so
 * one captured CUDA graph serves every step -- the host increments *step_dev before each replay). */
int mmvae_adam_step(int64_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                    float lr, float beta1, float beta2, float eps, float weight_decay,
                    int64_t step, const int64_t* step_dev, float grad_scale, void* stream);

/* i-th Conv2d / ConvTranspose2d in execution order: name ("encoder.layer1.0.conv1") and
 * shape = {kind (0 conv, 1 transposed), k, stride, padding, Ci, Co, H_in, H_out}. */
int mmvae_conv_entry(const mmvae_desc* d, int32_t i, char* name, size_t name_cap, int32_t shape[8]);

/* Measurement hook: enqueue ONE launch of the production tcgen05 kernel of conv `conv_index` (mmvae_conv_entry order)
 * in direction `dir` (0 forward + fused BatchNorm statistics, 1 data gradient, 2 weight gradient) on whatever the
 * workspace currently holds (run a forward / backward first).  algo_bytes / algo_flops receive the algorithmic
 * traffic (bf16 input + output + weights, each touched once) and 2*MAC of that launch, so that a caller can bracket
 * it with events and report achieved GB/s or FLOP/s (bench.py's roofline).  grads_scratch: n_params floats (dir 2). */
int mmvae_bench_conv(const mmvae_desc* d, int32_t conv_index, int32_t dir, const float* params, void* workspace,
                     size_t workspace_bytes, float* grads_scratch, int64_t* algo_bytes, int64_t* algo_flops, void* stream);

/* Measurement hook of the notebook variant's dominant kernels (nb_tail.cu, decoder.conv4 on 128-pixel rows): enqueue ONE
 * launch of `which` = 0 forward fused with the softmax cross-entropy (needs target), 1 data gradient, 2 weight + bias
 * gradient, on whatever the workspace holds after a mmvae_forward / mmvae_nb_loss_backward pair.  algo_bytes = bf16
 * tensors the launch must touch once (input rows, d logits, targets, weights), algo_flops = 2*MAC. */
int mmvae_nb_bench_tail(const mmvae_desc* d, int32_t which, const float* params, const int64_t* target, void* workspace,
                        size_t workspace_bytes, float* grads_scratch, int64_t* algo_bytes, int64_t* algo_flops, void* stream);

/* Debugging hook: device buffer of [444 CTAs][16] uint64 that the tcgen05 conv kernel fills with %globaltimer stamps
 * of its pipeline milestones (scripts/trace_conv.py prints the timeline); NULL (the default) turns tracing off. */
void mmvae_debug_set_trace(void* device_buffer);

/* Self-test of the tcgen05 kernels: every conv of the model that the tensor-core path covers is run in
 * all three directions (forward + BatchNorm statistics, weight gradient, data gradient) through both
 * the tcgen05 kernel and the fp32-FMA SIMT kernel on identical pseudo-random bf16 inputs.
 * report[16*i + 4*j + {0,1,2}] = {sum (tc - simt)^2, sum simt^2, max |tc - simt|} for conv i and
 * j = 0 forward output, 1 BatchNorm (mean, rstd), 2 weight gradient, 3 data gradient (zeros = not run).
 * grads_a / grads_b: two scratch gradient arenas (n_params floats each).  Overwrites the workspace. */
int mmvae_selftest_tc(const mmvae_desc* d, const float* params, void* workspace, size_t workspace_bytes,
                      float* grads_a, float* grads_b, float* report, int32_t report_cap, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMVAE_H_ */
