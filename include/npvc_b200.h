/*
 * npvc_b200.h -- C-ABI of the B200-native ConvVAE hot path (libnpvc_b200.so).
 *
 * The reference (JeremyCCHsu/vae-npvc) has NO native interface: the path is TensorFlow-1.2 graph
 * ops issued from Python.  Each entry point below replaces the group of TF ops the cited
 * reference lines issue; `model/vae.py` / `trainer/vae.py` in this repo bind them with ctypes
 * behind the reference's own ConvVAE / VAETrainer plugin surface.
 *
 * Conventions
 *   - plain C types only; every pointer named d_* is a DEVICE pointer to contiguous fp32 (labels:
 *     int64), owned by the caller.  The library owns only immutable plan tables (index maps built
 *     from the architecture), uploaded on the first device call and freed by npvc_destroy().
 *   - every call takes the CUDA stream (cudaStream_t passed as void*) to enqueue on; nothing
 *     synchronises the device; nothing runs on the legacy default stream unless stream == NULL.
 *     npvc_loss_fwd_bwd may fork part of its backward pass (the weight gradients) onto an internal
 *     non-blocking stream; that work is ordered after / before the caller's stream by events, so the
 *     call keeps plain stream semantics (and can be captured in a CUDA graph).
 *   - return value: 0 = ok, non-zero = error; npvc_last_error() gives the thread-local message.
 *   - there is NO CPU fallback: with no usable GPU every compute entry returns NPVC_ERR_CUDA.
 *   - parameters travel as ONE flat fp32 buffer `theta` = concatenation of the TF variables in
 *     tf.trainable_variables() creation order, each in its TF layout (npvc_param_table()).
 *     Gradients / Adam moments use the same flat layout (one NCCL all-reduce bucket).
 *   - frames are frame-major: x[n,513] == NCHW [n,1,513,1] (analyzer.py:116-122).
 *   - labels must lie in [0, y_dim); they are not range-checked on the device.  A label outside the range never causes
 *     an out-of-bounds access, but it selects no row of the per-speaker table, which carries the speaker term
 *     emb[y].W_y AND the three merge biases (model/vae.py:51-61 folded per speaker): such a frame is decoded without
 *     them, unlike tf.nn.embedding_lookup on a GPU (zero embedding row, biases still added).
 */
#ifndef NPVC_B200_H
#define NPVC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NPVC_MAX_LAYERS 8

enum {
  NPVC_OK = 0,
  NPVC_ERR_ARG = 1,      /* bad argument / unsupported architecture */
  NPVC_ERR_CUDA = 2,     /* CUDA runtime error or no device */
  NPVC_ERR_WORKSPACE = 3 /* workspace too small */
};

/* Architecture = the consumed keys of architecture-*.json (architecture-vae-vcc2016.json:2-20). */
typedef struct npvc_arch {
  int32_t in_h;                        /* hwc[0] = 513 */
  int32_t z_dim;                       /* 128 */
  int32_t y_dim;                       /* 10 */
  int32_t n_enc;                       /* len(encoder.output) */
  int32_t enc_out[NPVC_MAX_LAYERS];    /* encoder.output */
  int32_t enc_kernel[NPVC_MAX_LAYERS]; /* encoder.kernel[i][0] */
  int32_t enc_stride[NPVC_MAX_LAYERS]; /* encoder.stride[i][0] */
  int32_t gen_h, gen_c;                /* generator.hwc = [h, 1, c] */
  int32_t n_gen;
  int32_t gen_out[NPVC_MAX_LAYERS];
  int32_t gen_kernel[NPVC_MAX_LAYERS];
  int32_t gen_stride[NPVC_MAX_LAYERS];
} npvc_arch;

typedef struct npvc_handle npvc_handle;

/* One TF variable of the path (SURVEY 8a "names"). */
typedef struct npvc_param_desc {
  char name[96];      /* e.g. "Encoder/Conv2d-0/Conv2d-0/kernel" */
  int64_t offset;     /* float offset into theta */
  int64_t size;       /* number of floats */
  int32_t rank;
  int32_t shape[4];   /* TF shape */
  int32_t fan_in;     /* glorot fans (0 for biases / LN params) */
  int32_t fan_out;
  int32_t init;       /* 0 = glorot_uniform, 1 = zeros, 2 = ones */
} npvc_param_desc;

/* ---- lifetime / introspection (host only, usable without a GPU) ------------------------- */

/* Replaces ConvVAE.__init__ + _sanity_check (model/vae.py:9-39): validates the architecture and
 * builds the launch plan.  max_chunk = frames processed per internal pass (0 = default 16384). */
int npvc_create(const npvc_arch* arch, int64_t max_chunk, npvc_handle** out);
void npvc_destroy(npvc_handle* h);
const char* npvc_last_error(void);
const char* npvc_version(void);

int64_t npvc_param_count(const npvc_handle* h);          /* 939,162 for the VCC2016 arch */
int32_t npvc_param_tensors(const npvc_handle* h);        /* 44 */
int npvc_param_table(const npvc_handle* h, npvc_param_desc* out, int32_t max);

/* Bytes of caller-owned device workspace for a call over n frames.
 * train != 0: loss_fwd_bwd (activations kept for backward); 0: encode/decode only. */
int64_t npvc_workspace_bytes(const npvc_handle* h, int64_t n_frames, int32_t train);

/* Launch plan as JSON (buffers, ops, views) -- used by the CPU plan-interpreter tests. */
const char* npvc_plan_json(const npvc_handle* h);
/* Copy a named int32 plan table ("pack_src", "pack16_src", "pack_list", "unpack_ptr", "unpack_idx") to host memory.
 * Returns its length; copies at most `max` entries when out != NULL. */
int64_t npvc_plan_table(const npvc_handle* h, const char* name, int32_t* out, int64_t max);
/* Number of kernels this library has launched through this handle (bench: gpu_launches). */
int64_t npvc_launch_count(const npvc_handle* h);

/* ---- device entry points ---------------------------------------------------------------- */

/* Re-layout theta into GEMM-ready operand matrices inside the workspace.  Must be called after
 * theta changes and before encode/decode (loss_fwd_bwd does it itself when repack != 0). */
int npvc_pack_weights(npvc_handle* h, const float* d_theta, void* d_ws, int64_t ws_bytes, void* stream);

/* ConvVAE.encode / _encoder (model/vae.py:72-82,139-141): x[n,513] -> mu[n,z], lv[n,z]. */
int npvc_encode(npvc_handle* h, const float* d_theta, const float* d_x, int64_t n,
                float* d_mu, float* d_lv, void* d_ws, int64_t ws_bytes, void* stream);

/* GaussianSampleLayer (util/layers.py:152-156) with the N(0,1) draw explicit:
 * z = mu + eps * sqrt(exp(lv)). */
int npvc_sample(npvc_handle* h, const float* d_mu, const float* d_lv, const float* d_eps,
                int64_t n, float* d_z, void* stream);

/* ConvVAE.decode / _generator (model/vae.py:84-103,143-145): z[n,z], y[n] -> xh[n,513]
 * (== NHWC [n,513,1,1]; the nchw_to_nhwc of util/image.py:4-5 is a relabel when C=W=1). */
int npvc_decode(npvc_handle* h, const float* d_theta, const float* d_z, const int64_t* d_y,
                int64_t n, float* d_xh, void* d_ws, int64_t ws_bytes, void* stream);

/* ConvVAE.loss (model/vae.py:106-130) + the autodiff of optimizer.minimize (trainer/vae.py:24):
 * forward, KL + Gaussian log-density, backward.  Outputs z, mu, lv [n,z], xh [n,513],
 * d_losses[3] = {G, D_KL, logP}, d_grad[param_count] = d G / d theta (overwritten).
 * Any of d_z/d_mu/d_lv/d_xh may be NULL.  d_grad == NULL: forward + losses only.
 * The loss means (and the gradient) are taken over the n frames of this call; data-parallel ranks
 * average their gradients afterwards (all-reduce SUM, then grad_scale = 1/world in npvc_adam_step). */
int npvc_loss_fwd_bwd(npvc_handle* h, const float* d_theta, const float* d_x, const int64_t* d_y,
                      const float* d_eps, int64_t n, float* d_z, float* d_mu, float* d_lv,
                      float* d_xh, float* d_losses, float* d_grad, int32_t repack,
                      void* d_ws, int64_t ws_bytes, void* stream);

/* Device-resident state of a training loop (32 bytes of caller-owned device memory, zero-initialised except `seed`):
 * what changes from step to step lives on the device, so the launch sequence of a whole training step is identical
 * every step and can be captured once in a CUDA graph and replayed.
 *   seed   key of the in-kernel N(0,1) sampler (Philox4x32-10 + Box-Muller)
 *   draws  number of completed npvc_train_fwd_bwd passes = third counter word of the sampler
 *   step   number of completed passes that produced a gradient = the t of the Adam update that follows */
typedef struct npvc_step_state { uint64_t seed; int64_t draws; int64_t step; int64_t reserved; } npvc_step_state;

/* npvc_loss_fwd_bwd with the tf.random_normal of GaussianSampleLayer (util/layers.py:154) drawn IN-KERNEL:
 * eps[frame, d] = Normal(Philox4x32-10(key = seed, counter = (d, frame_offset + frame, draws))) -- a function of
 * (seed, pass, frame, dim) only, independent of chunking; the backward regenerates it (no eps buffer exists).
 * frame_offset: index of this call's first frame in the job's batch (data-parallel ranks pass rank * n to draw
 * different noise from one seed).  On completion draws += 1 and, when d_grad != NULL, step += 1 (in stream order). */
int npvc_train_fwd_bwd(npvc_handle* h, const float* d_theta, const float* d_x, const int64_t* d_y,
                       npvc_step_state* d_state, int64_t frame_offset, int64_t n, float* d_z, float* d_mu,
                       float* d_lv, float* d_xh, float* d_losses, float* d_grad, int32_t repack,
                       void* d_ws, int64_t ws_bytes, void* stream);
/* The draw alone: d_eps[n, z] = what npvc_train_fwd_bwd would use for these frames at the state's current `draws`. */
int npvc_normal_draw(npvc_handle* h, const npvc_step_state* d_state, int64_t frame_offset, int64_t n,
                     float* d_eps, void* stream);

/* tf.train.AdamOptimizer.apply_gradients, TF form (trainer/vae.py:16-24):
 * lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v EMAs; theta -= lr_t*m/(sqrt(v)+eps).  grad is multiplied
 * by grad_scale first (1/world_size after an all-reduce SUM).  step t >= 1. */
int npvc_adam_step(npvc_handle* h, float* d_theta, const float* d_grad, float* d_m, float* d_v,
                   int64_t n_params, int64_t step, float lr, float beta1, float beta2, float eps,
                   float grad_scale, void* stream);

/* npvc_adam_step with t = d_state->step read on the device (see npvc_step_state). */
int npvc_adam_step_dev(npvc_handle* h, float* d_theta, const float* d_grad, float* d_m, float* d_v,
                       int64_t n_params, const npvc_step_state* d_state, float lr, float beta1, float beta2,
                       float eps, float grad_scale, void* stream);

/* Tanhize.forward_process / backward_process (analyzer.py:82-87) over [n, dim] with per-bin
 * xmin/xmax [dim].  In-place allowed. */
int npvc_tanhize_forward(npvc_handle* h, const float* d_x, const float* d_xmin, const float* d_xmax,
                         int64_t n, int32_t dim, float* d_out, void* stream);
int npvc_tanhize_backward(npvc_handle* h, const float* d_x, const float* d_xmin, const float* d_xmax,
                          int64_t n, int32_t dim, float* d_out, void* stream);

/* Frame-reader kernel (analyzer.py:111-135): records[n, rec_floats] (sp|ap|f0|en|spk) ->
 * x[n,513] Tanhize'd + y[n] int64 speaker, fused (feeds the pinned-memory loader). */
int npvc_unpack_records(npvc_handle* h, const float* d_records, int64_t n, int32_t rec_floats,
                        int32_t sp_dim, const float* d_xmin, const float* d_xmax,
                        float* d_x, int64_t* d_y, void* stream);

/* Per-op device timing (bench.py roofline): enable != 0 brackets every op of the following calls
 * with CUDA events on the launching stream; npvc_profile_json() synchronises those events and
 * returns [{"name","kind","calls","ms","rows","K","N"}...] aggregated per op, then clears them. */
int npvc_profile_enable(npvc_handle* h, int32_t enable);
const char* npvc_profile_json(npvc_handle* h);

/* Debug: copy a named workspace buffer of the LAST pass (first chunk) to a device pointer.
 * Returns its float count per frame (negative on error). */
int64_t npvc_debug_buffer(npvc_handle* h, const char* name, const void* d_ws, float* d_out,
                          int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NPVC_B200_H */
