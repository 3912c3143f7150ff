/*
 * b200dp.h — C ABI of the B200-native data-parallel gradient-reduction path.
 *
 * This is the drop-in boundary.  The reference (huggingface/accelerate) has no FFI of its own:
 * every byte of this path is executed by PyTorch's DistributedDataParallel / c10d::Reducer /
 * ProcessGroupNCCL / GradScaler / clip_grad_norm_, which the reference reaches through the Python
 * call sites cited on each entry point below (paths relative to /root/reference).  The functions
 * here are what a binding for that path would call; INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - plain C types only: pointers, sizes, ints.  Streams are `cudaStream_t` passed as void*.
 *   - every function returns 0 on success, a negative B200DP_E* code on failure;
 *     b200dp_last_error() returns a thread-local human-readable message for the last failure.
 *   - every data-path call is stream-ordered and host-non-blocking unless stated otherwise.
 *   - "segment" = a symmetric allocation: every rank allocates the same number of bytes and maps
 *     every peer's copy (NVLink P2P), so kernels address peers with plain ld/st.global.
 *   - collectives must be issued by all ranks in the same order per channel (same rule DDP has).
 */
#ifndef B200DP_H
#define B200DP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200DP_VERSION 100

#define B200DP_MAX_WORLD 16
#define B200DP_MAX_BLOCKS 304  /* >= 2 x 148 SMs */
#define B200DP_MAX_SLOTS 1024  /* norm/found-inf record slots (one per bucket) */
#define B200DP_SMALL_MAX 128   /* bytes per rank in b200dp_exchange_small */
#define B200DP_MT_MAX 640      /* tensors per launch in the multi-tensor entry points */

/* error codes */
#define B200DP_OK 0
#define B200DP_ECUDA -1    /* CUDA runtime / driver call failed */
#define B200DP_EINVAL -2   /* bad argument */
#define B200DP_ESTATE -3   /* call sequence violated (e.g. segment not mapped) */
#define B200DP_ETIMEOUT -4 /* a device-side flag wait or host wait timed out (peer hang) */
#define B200DP_ENOSUP -5   /* not supported on this system (e.g. multicast) */
#define B200DP_ESYS -6     /* OS error (socket / fd passing) */

/* element types */
#define B200DP_F32 0
#define B200DP_BF16 1
#define B200DP_F16 2
#define B200DP_F64 3
#define B200DP_I32 4
#define B200DP_I64 5
#define B200DP_U8 6

/* segment handle kinds (b200dp_seg_export) */
#define B200DP_HANDLE_CUDAIPC 0 /* 64-byte cudaIpcMemHandle_t, cudaMalloc memory */
#define B200DP_HANDLE_POSIXFD 1 /* cuMemCreate allocation exported as a POSIX fd (multicast-capable) */

/* channels: independent flag pads + sequence counters; kernels on one channel must be stream-ordered */
#define B200DP_CH_GRAD 0  /* gradient buckets (reducer's side stream) */
#define B200DP_CH_USER 1  /* gather / reduce / broadcast on the caller's stream */
#define B200DP_CH_SMALL 2 /* b200dp_exchange_small */
#define B200DP_CH_PUSH 3  /* b200dp_push (its window parity is derived from this channel's counters) */
#define B200DP_NUM_CHANNELS 4

/* allreduce algorithm selectors */
#define B200DP_ALGO_AUTO 0
#define B200DP_ALGO_TWOSHOT 1 /* reduce-scatter by peer loads, all-gather by peer stores */
#define B200DP_ALGO_NVLS 2    /* multimem.ld_reduce + multimem.st on the multicast mapping */
#define B200DP_ALGO_ONESHOT 3 /* out-of-place only: every rank loads all N copies */

/* "no stream" sentinel for optional stream arguments.  NULL cannot be used: cudaStream_t 0 is the
 * (legacy) default stream, which is a perfectly valid stream to order after. */
#define B200DP_NO_STREAM ((void*)(intptr_t)-1)

/* flags for b200dp_allreduce_bucket */
#define B200DP_F_STATS 1 /* accumulate sum-of-squares / max-abs / non-finite into the slot record */
#define B200DP_F_NO_END_BARRIER 2 /* skip the closing cross-rank barrier: legal when another kernel of the
                                   * same channel/stream follows (its opening barrier proves the peers finished
                                   * this one); the LAST kernel of a sequence must be followed by b200dp_fence */

/* b200dp_query keys */
#define B200DP_Q_LAUNCHES 0      /* kernels launched by this context so far */
#define B200DP_Q_MULTICAST_OK 1  /* device + driver expose multicast objects */
#define B200DP_Q_VMM_OK 2        /* cuMemCreate + POSIX-fd export usable */
#define B200DP_Q_STATUS 3        /* device-side status word (0 = ok, else B200DP_E*) */
#define B200DP_Q_SM_COUNT 4
#define B200DP_Q_COMM_STREAM 5   /* cudaStream_t of the context's high-priority side stream */

typedef struct b200dp_ctx b200dp_ctx;

int b200dp_version(void);
const char* b200dp_last_error(void);

/* ---- context --------------------------------------------------------------------------------
 * Replaces: the per-process NCCL communicator a DDP wrapper owns after
 * torch.distributed.init_process_group (src/accelerate/state.py:244) — only for data movement;
 * rendezvous and barriers stay with torch.distributed.  One context per process/GPU. */
int b200dp_init(int rank, int world, int device, b200dp_ctx** out);
int b200dp_finalize(b200dp_ctx* ctx);
int b200dp_query(b200dp_ctx* ctx, int key, long long* out);
/* timeout for device-side flag waits, milliseconds (default 600000 = the reference's NCCL
 * watchdog default, src/accelerate/utils/dataclasses.py:292-299). */
int b200dp_set_timeout_ms(b200dp_ctx* ctx, long long ms);
/* clear a latched device-side error (tests). */
int b200dp_reset_status(b200dp_ctx* ctx);

/* ---- symmetric segments ---------------------------------------------------------------------
 * Replaces: the flat bucket tensors c10d::Reducer allocates (torch reducer.hpp, reached from
 * src/accelerate/accelerator.py:1892-1894) and NCCL's internal registered buffers. */
int b200dp_seg_create(b200dp_ctx* ctx, size_t bytes, int prefer_kind, int* seg, void** local_ptr,
                      size_t* padded_bytes, int* kind);
int b200dp_seg_export(b200dp_ctx* ctx, int seg, unsigned char handle[64], int* fd);
int b200dp_seg_import(b200dp_ctx* ctx, int seg, int peer, int kind, const unsigned char handle[64],
                      int fd);
int b200dp_seg_peer_ptr(b200dp_ctx* ctx, int seg, int peer, void** ptr);
int b200dp_seg_destroy(b200dp_ctx* ctx, int seg);
/* multicast (NVLS) binding of a POSIXFD segment: rank 0 creates + exports, all attach, barrier
 * (caller), all bind.  After bind, b200dp_seg_mc_ptr returns the multicast VA. */
int b200dp_mc_create(b200dp_ctx* ctx, int seg, int* fd);
int b200dp_mc_attach(b200dp_ctx* ctx, int seg, int fd);
int b200dp_mc_bind(b200dp_ctx* ctx, int seg);
int b200dp_seg_mc_ptr(b200dp_ctx* ctx, int seg, void** ptr);
/* declare the (already mapped) segment that holds flag pads, slot records and small-exchange
 * slots; must be >= b200dp_control_bytes() and zero-initialised (seg_create zeroes it). */
size_t b200dp_control_bytes(void);
int b200dp_set_control(b200dp_ctx* ctx, int seg);

/* ---- fd passing between the ranks of one node (SCM_RIGHTS over abstract unix datagram sockets) */
int b200dp_fdx_open(b200dp_ctx* ctx, const char* name);
int b200dp_fdx_send(b200dp_ctx* ctx, const char* peer_name, int fd, int tag);
int b200dp_fdx_recv(b200dp_ctx* ctx, int* fd, int* tag, int timeout_ms);
int b200dp_fdx_close(b200dp_ctx* ctx);

/* ---- stream plumbing ------------------------------------------------------------------------ */
/* make `waiter` wait for everything enqueued so far on `signaller` (event record + wait). */
int b200dp_stream_wait(b200dp_ctx* ctx, void* waiter, void* signaller);

/* ---- K1: fused gradient-bucket allreduce ----------------------------------------------------
 * Replaces: Reducer::mark_variable_ready_dense pre-divide + ncclAllReduce(SUM) per bucket
 * (torch reducer.hpp:276,524; configured by DistributedDataParallelKwargs,
 * src/accelerate/utils/dataclasses.py:154-237) and, with B200DP_F_STATS, the first read pass of
 * torch clip_grad_norm_ / GradScaler.unscale_'s non-finite check
 * (src/accelerate/accelerator.py:2911-3007).
 * In place on segment `seg`, elements [elem_off, elem_off+count) of type dtype:
 *   x_i <- sum_{r=0..N-1} ( x_i^{(r)} * prescale )   in rank order, fp32 accumulation,
 * identical bits on every rank.  Stats of the reduced values go to record `slot`.
 * If after_stream != B200DP_NO_STREAM the launch is ordered after everything already enqueued on
 * after_stream (event record + wait; after_stream may be 0 = the default stream). */
int b200dp_allreduce_bucket(b200dp_ctx* ctx, int seg, size_t elem_off, size_t count, int dtype,
                            float prescale, int slot, int flags, int algo, int channel,
                            int max_blocks, void* after_stream, void* stream);

/* one-CTA cross-rank barrier on `channel`: when it retires on this rank, every peer has finished every
 * earlier kernel of that channel (so all their stores into this rank's memory have landed).  Closes a
 * sequence of B200DP_F_NO_END_BARRIER launches (the reducer issues it once per backward). */
int b200dp_fence(b200dp_ctx* ctx, int channel, void* stream);

/* out-of-place one-shot: dst_i = sum_r staging_r[i] (rank order) then / divisor (true division
 * for floating types; divisor 1 = plain sum).  Replaces operations.reduce's
 * all_reduce(SUM) + `/= N` (src/accelerate/utils/operations.py:871-873) for small tensors. */
int b200dp_allreduce_oneshot(b200dp_ctx* ctx, int seg, size_t byte_off, size_t count, int dtype,
                             double divisor, void* dst, int channel, int max_blocks, void* stream);

/* ---- K6: clip / unscale epilogue ------------------------------------------------------------
 * Replaces: torch.nn.utils.clip_grad_norm_ (src/accelerate/accelerator.py:3007) and
 * GradScaler._unscale_grads_ (:2944).  finalize: total = norm over records [slot0, slot0+nslots)
 * x all ranks (norm_type 2.0 or +inf), times inv_scale if scale_ptr != NULL
 * (inv_scale = (float)(1.0/(double)*scale_ptr));
 * out[0]=total_norm, out[1]=multiplier = inv_scale * min(1, max_norm/(total+1e-6)) (max_norm<0:
 * no clipping), out[2]=found_inf (0/1).  `out` is device memory (3 floats). */
int b200dp_finalize_norm(b200dp_ctx* ctx, int slot0, int nslots, int all_ranks, double norm_type,
                         float max_norm, const float* scale_ptr, float* out, void* stream);
/* x *= *mult_ptr over a flat range (no traffic at all when *mult_ptr == 1). */
int b200dp_scale_flat(b200dp_ctx* ctx, void* ptr, size_t count, int dtype, const float* mult_ptr,
                      int max_blocks, void* stream);
/* clamp to [-v, v] (clip_grad_value_, src/accelerate/accelerator.py:3034). */
int b200dp_clamp_flat(b200dp_ctx* ctx, void* ptr, size_t count, int dtype, float v, int max_blocks,
                      void* stream);

/* dst = (dst_dtype)src, then / divisor in dst_dtype (divisor 1 = plain cast); f32 <-> bf16/f16 only.
 * The half-precision wire format of the gradient allreduce (DDPCommunicationHookType.FP16/BF16,
 * src/accelerate/utils/dataclasses.py:206-216 -> torch default_hooks fp16/bf16_compress_hook). */
int b200dp_cast(b200dp_ctx* ctx, const void* src, int src_dtype, void* dst, int dst_dtype, size_t count,
                float divisor, int max_blocks, void* stream);

/* ---- K5: multi-tensor variants for gradients that do not live in a segment (world_size 1,
 * parameter subsets).  ptrs/numels are host arrays of length n; same dtype for all.
 * multi_stats writes ceil(n / B200DP_MT_MAX) consecutive local records starting at `slot`
 * (one per launch); pass the same slot range to b200dp_finalize_norm with all_ranks=0. */
int b200dp_multi_stats(b200dp_ctx* ctx, const void* const* ptrs, const size_t* numels, int n,
                       int dtype, int slot, void* stream);
int b200dp_multi_scale(b200dp_ctx* ctx, void* const* ptrs, const size_t* numels, int n, int dtype,
                       const float* mult_ptr, void* stream);

/* ---- K9: fused optimizer consumption -------------------------------------
 * Replaces: the ~10 foreach launches of torch.optim.Adam/AdamW.step that AcceleratedOptimizer.step reaches
 * (src/accelerate/optimizer.py:145-181) plus the clip multiply pass, for fp32 parameters whose gradients and
 * moments live in two flat buffers: one pass reads p, g, m, v and writes p, m, v.  params[i] / grads[i] have
 * numels[i] elements (grads[i] is a view of the fusion buffer, or any fp32 device tensor); its moments start at
 * element offs[i] of exp_avg / exp_avg_sq.  Hyper-parameters as torch computes them on the host for step t:
 *   decay_mul = 1 - lr*wd (AdamW; pass 1 for none), l2 = wd (Adam's coupled L2; 0 for none),
 *   neg_step_size = -lr/(1-beta1^t), bc2_sqrt = sqrt(1-beta2^t).
 * They are doubles because torch derives 1-beta1 and 1-beta2 in double before rounding to float
 * (float(1 - 0.999) and 1.0f - 0.999f differ by 5e-5 relative).
 * mult_ptr (device, may be NULL): gradient multiplier (clip coefficient x 1/loss-scale); found_inf_ptr (device,
 * may be NULL): non-zero skips the whole step. */
int b200dp_multi_adam(b200dp_ctx* ctx, void* const* params, const void* const* grads, const size_t* numels,
                      const size_t* offs, int n, void* exp_avg, void* exp_avg_sq, double decay_mul, double l2,
                      double beta1, double beta2, double neg_step_size, double bc2_sqrt, double eps,
                      const float* mult_ptr, const float* found_inf_ptr, void* stream);

/* ---- K4: bucket pack / unpack (multi-tensor copy) -------------------------------------------
 * Replaces: the per-parameter grad->bucket and bucket->grad copy kernels of the Reducer that
 * src/accelerate/accelerator.py:1892-1894 installs (torch reducer.hpp:276,286,328).  n (src,dst,bytes) triples, one launch.  use_bulk=1 moves
 * 16-byte-aligned spans with cp.async.bulk (TMA engine) through shared memory. */
int b200dp_pack(b200dp_ctx* ctx, const void* const* srcs, void* const* dsts, const size_t* nbytes,
                int n, int use_bulk, void* stream);

/* ---- K2/K3: allgather / broadcast by peer loads ---------------------------------------------
 * Replaces: all_gather_into_tensor (src/accelerate/utils/operations.py:342) and
 * torch.distributed.broadcast (:522).  Every rank has already placed `nbytes` at byte_off of its
 * copy of `seg` (b200dp_pack); this pulls rank r's bytes to dst + (r-first)*dst_stride for
 * r in [first, first+nranks).  Bit-exact. */
int b200dp_pull(b200dp_ctx* ctx, int seg, size_t byte_off, size_t nbytes, void* dst,
                size_t dst_stride, int first, int nranks, int channel, int max_blocks, void* stream);

/* ---- K2c: fused large-message allgather: stage + pull in ONE launch -----------------------------
 * Replaces: torch.distributed.all_gather_into_tensor on a tensor that lives in ordinary torch memory
 * (src/accelerate/utils/operations.py:337-343).  CTA b copies ITS portion of `src` into this rank's copy of
 * `seg` at byte_off, meets CTA b of every rank at a flag barrier and pulls that portion from every rank into
 * dst + r*dst_stride: no separate pack launch, no grid-wide barrier.  nbytes <= capacity of the window. */
int b200dp_gather(b200dp_ctx* ctx, int seg, size_t byte_off, const void* src, size_t nbytes, void* dst,
                  size_t dst_stride, int channel, int max_blocks, void* stream);

/* ---- K1c: fused large-message out-of-place allreduce: stage + reduce-scatter + all-gather + copy-out in ONE
 * launch.  Replaces: clone + torch.distributed.all_reduce(SUM) (+ `/= N`) (operations.py:859-874).
 * dst = sum_r src_r (rank order, fp32 accumulate; NVLS: switch order), then T(float(T(sum)) / divisor) if
 * divisor != 1.  src is never written.  f32 / bf16 / f16; algo TWOSHOT, NVLS or AUTO. */
int b200dp_reduce(b200dp_ctx* ctx, int seg, size_t byte_off, const void* src, size_t count, int dtype,
                  double divisor, void* dst, int algo, int channel, int max_blocks, void* stream);

/* Copy-engine read of `nbytes` from rank `peer`'s copy of `seg` into local memory (cudaMemcpyAsync over the peer
 * mapping): the NVLink yardstick bench.py measures the kernels against.  Not used on the data path. */
int b200dp_peer_copy(b200dp_ctx* ctx, int seg, int peer, size_t byte_off, void* dst, size_t nbytes, void* stream);

/* ---- K2b: latency path (one launch, one barrier) for messages up to b200dp_push_max_bytes() per rank.
 * Replaces all_gather_into_tensor / all_reduce for small tensors (operations.py:342,871).  The last
 * b200dp_push_area_bytes() bytes of segment `seg` are used as per-CTA, double-buffered windows, so
 * [0, seg_bytes - b200dp_push_area_bytes()) stays free for b200dp_pull / b200dp_allreduce_*.
 * mode 0: dst[r*nbytes : (r+1)*nbytes] = rank r's src (any dtype; dtype ignored).
 * mode 1: dst = sum_r src_r in rank order, then / divisor (nbytes = count * sizeof(dtype)). */
size_t b200dp_push_max_bytes(void);
size_t b200dp_push_area_bytes(void);
int b200dp_push(b200dp_ctx* ctx, int seg, int mode, const void* src, size_t nbytes, int dtype, double divisor,
                void* dst, int channel, void* stream);

/* ---- K8: tiny host-visible exchange ---------------------------------------------------------
 * Replaces: the shape all-gather + .cpu() sync of pad_across_processes
 * (src/accelerate/utils/operations.py:774-775).  nbytes <= B200DP_SMALL_MAX from host memory;
 * host_out receives world*nbytes.  BLOCKS the host until the data is there (that is the op). */
int b200dp_exchange_small(b200dp_ctx* ctx, const void* host_src, int nbytes, void* host_out,
                          void* stream);

/* ---- K7: fused pad copy ---------------------------------------------------------------------
 * Replaces: new_zeros + pad_index, slice-assign (src/accelerate/utils/operations.py:784-791).
 * src viewed as [outer, old_dim, inner_bytes], dst as [outer, new_dim, inner_bytes];
 * data goes to the head (pad_first=0) or tail (pad_first=1) of new_dim; the rest is filled
 * with the elem_size-byte pattern in `pad_bits`. */
int b200dp_pad_copy(b200dp_ctx* ctx, const void* src, void* dst, size_t outer, size_t old_dim,
                    size_t new_dim, size_t inner_bytes, int elem_size, uint64_t pad_bits,
                    int pad_first, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200DP_H */
