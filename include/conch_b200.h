/*
 * conch_b200.h -- C ABI of the B200-native mixed-precision GEMM library (libconch_b200.so).
 *
 * This is the drop-in boundary for ONE reference path:
 *   conch.ops.quantization.gemm.mixed_precision_gemm      (reference conch/ops/quantization/gemm.py:133-168)
 *     -> mixed_precision_gemm_launcher                     (reference conch/kernels/quantization/gemm.py:482-545)
 *     -> Triton _gemm_kernel + _dequantize                 (reference conch/kernels/quantization/gemm.py:176-457)
 * Every entry point below replaces the Triton launch `_gemm_kernel[grid](...)` at
 * conch/kernels/quantization/gemm.py:500-545 (or, for cb200_dequantize, the `_dequantize` stage at :176-216
 * run on its own as a test aid).  The host side that binds it is conch_b200/_lib.py (ctypes); the stub a
 * maintainer of the reference would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C: pointers + sizes only, no torch / C++ types.
 *   - every pointer is a DEVICE pointer unless the parameter name ends in `_host`.
 *   - the caller owns every buffer (inputs, `out`, `workspace`); nothing is allocated or freed here,
 *     nothing synchronises; all work is enqueued on `stream` (a cudaStream_t passed as void*), so every
 *     call is CUDA-graph capturable.
 *   - return value: 0 = success; < 0 = cb200_status (argument / unsupported); > 0 = cudaError_t.
 *     cb200_last_error() returns a thread-local human readable message for the last non-zero return.
 *   - tensors are dense row-major:
 *        x            [M, K]      fp16 | bf16                      (reference arg `x`)
 *        w_q_packed   [K/pf, N]   int32, pf = 32/bits, value k = pf*row + i sits in bits [bits*i, bits*(i+1))
 *                                                                  (reference conch/third_party/vllm/quant_utils.py:104-125)
 *        w_s          [K/g, N]    same dtype as x                  (reference arg `w_s`)
 *        w_zp         [K/g, N]    int32 | uint8 | dtype of x, or 1 element, or NULL
 *        out          [M, N]      fp16 | bf16 | fp32
 *     W[k,n] = RN_act( (q[k,n] - weight_bias - zp[k/g,n]) * s[k/g,n] ),  out = cast(sum_k fp32(x[m,k] * W[k,n]))
 *                                                                  (reference conch/kernels/quantization/gemm.py:192-216, 401, 443-457)
 */
#ifndef CONCH_B200_H_
#define CONCH_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB200_VERSION 100 /* 1.0.0 */

#if defined(__GNUC__)
#define CB200_API __attribute__((visibility("default")))
#else
#define CB200_API
#endif

/* element types (activations, scales, outputs) */
enum cb200_dtype { CB200_F16 = 0, CB200_BF16 = 1, CB200_F32 = 2 };

/* 8-bit operand formats of the scaled-GEMM sibling path (reference: torch.int8 / torch.float8_e4m3fn) */
enum cb200_in_kind { CB200_IN_INT8 = 0, CB200_IN_FP8_E4M3 = 1 };

/* what `w_zp` points at (reference: w_zp None / int32 from quantize_weights / uint8 from vLLM / meta dtype) */
enum cb200_zp_kind { CB200_ZP_NONE = 0, CB200_ZP_I32 = 1, CB200_ZP_U8 = 2, CB200_ZP_ACT = 3 };

/* kernel families; 0 lets the library choose by regime */
enum cb200_variant {
  CB200_AUTO = 0,
  CB200_DECODE_MMA = 1,   /* <= 16 rows per pass (AUTO: M <= 8 where CB200_DECODE_STREAM has no stage layout for the problem): split-K skinny GEMM, 128-bit weight loads, in-register unpack, HMMA dot, last-arriver reduce */
  CB200_TC_SS = 2,        /* tcgen05: TMA -> smem dequant (K-major, 128B swizzle) -> UMMA (both operands smem), TMEM accumulators */
  CB200_TC_TS = 3,        /* tcgen05: TMA -> dequant straight into TMEM (W is the TMEM A operand), x tile from smem */
  CB200_TC2_TS = 4,       /* persistent tcgen05 kernel: m_tile 128, two TMEM accumulators, epilogue overlaps the next tile */
  CB200_DECODE_STREAM = 5 /* <= 16 rows per pass (AUTO: M <= 8, and M <= 16 for 4-bit weights): warp-specialised streaming GEMV -- a
                             producer warp keeps a ring of TMA weight tiles in shared memory full (it never waits for the stream order: under programmatic dependent launch
                             the ring fills while the previous kernel drains), 12-16 consumer warps turn nibbles into subnormal fp16
                             MMA operands, K is split over a thread-block cluster whose partial sums meet in distributed shared
                             memory (no workspace, no atomics, deterministic).  bf16 x int8 runs as two nibble planes.  Group
                             sizes that straddle a 16-packed-row stage (e.g. int4 groups of 192) are left to CB200_DECODE_MMA. */
};

enum cb200_status {
  CB200_OK = 0,
  CB200_ERR_ARG = -1,         /* null / misaligned pointer, non-positive size             -> ValueError        */
  CB200_ERR_SHAPE = -2,       /* K, N or group_size outside the supported lattice         -> ValueError        */
  CB200_ERR_UNSUPPORTED = -3, /* bits not in {4,8}, dtype combination not implemented     -> NotImplementedError */
  CB200_ERR_WORKSPACE = -4,   /* workspace too small for the chosen variant               -> ValueError        */
  CB200_ERR_DEVICE = -5       /* not an sm_100 device / driver entry point missing        -> RuntimeError      */
};

CB200_API int cb200_version(void);
CB200_API const char* cb200_last_error(void);

/* Bytes of scratch cb200_mpgemm may need for this problem (split-K partial sums + arrival counters).
 * The first CB200_WORKSPACE_COUNTER_BYTES bytes hold counters and MUST be zero before the first use;
 * kernels leave them zero again on exit, so the buffer can be reused forever without a memset. */
#define CB200_WORKSPACE_COUNTER_BYTES 65536
CB200_API size_t cb200_mpgemm_workspace_bytes(int64_t M, int64_t K, int64_t N, int weight_size_bits, int variant);

/* out[M,N] = x[M,K] @ dequant(w_q_packed, w_s, w_zp).
 * Replaces: mixed_precision_gemm_launcher(output, x, w_q_packed, w_s, w_zp, metadata)
 *           (reference conch/kernels/quantization/gemm.py:482-545).
 * group_size: elements of K sharing one scale row; -1 or K = per-channel.  Must be a multiple of 64 (or K),
 * like the reference kernel (its BK=64 K-block picks ONE scale row, gemm.py:329,354).  K % 64 == 0, N % 64 == 0. */
CB200_API int cb200_mpgemm(const void* x, const void* w_q_packed, const void* w_s, const void* w_zp, void* out,
                 int64_t M, int64_t K, int64_t N,
                 int weight_size_bits, int weight_bias, int64_t group_size,
                 int act_dtype, int out_dtype, int zp_kind, int zp_is_scalar,
                 void* workspace, size_t workspace_bytes,
                 int variant, void* stream);

/* cb200_mpgemm with a fused bias: out[m, n] = cast(sum_k x[m,k] W[k,n] + bias[n]), bias [N] in the OUTPUT dtype, added to the
 * fp32 sum before the single rounding (bias == NULL: plain cb200_mpgemm).
 * Replaces: the `output.add_(bias)` a caller of the reference op runs as a second kernel after mixed_precision_gemm_launcher
 * (the reference's own sibling op does exactly that: conch/ops/quantization/gemm.py:249-250). */
CB200_API int cb200_mpgemm_bias(const void* x, const void* w_q_packed, const void* w_s, const void* w_zp, const void* bias, void* out,
                 int64_t M, int64_t K, int64_t N,
                 int weight_size_bits, int weight_bias, int64_t group_size,
                 int act_dtype, int out_dtype, int zp_kind, int zp_is_scalar,
                 void* workspace, size_t workspace_bytes,
                 int variant, void* stream);

/* Multi-GPU epilogue (no counterpart in the reference, which has no distributed code): cb200_mpgemm whose
 * kernel ALSO stores the result into `num_peers` (<= 7) peer buffers -- peer_out[i] is a device pointer, valid in this
 * process, to an [M, N] buffer of another GPU (CUDA IPC / symmetric memory): typically this rank's slot of every other
 * rank's gather buffer, so an N-sharded layer needs no separate all-gather.  The stores are plain P2P writes over NVLink,
 * visible to the peer once this kernel has completed; synchronising the ranks (one barrier per token, say) is the
 * caller's job.  Decode kernels only: AUTO picks one for M <= 8 (M <= 16 where the streaming kernel supports the problem), variant = CB200_DECODE_MMA / CB200_DECODE_STREAM forces
 * it for any M (16 rows per pass); any other choice returns CB200_ERR_UNSUPPORTED -- large outputs are bandwidth-bound and
 * belong to an all-gather. */
CB200_API int cb200_mpgemm_push(const void* x, const void* w_q_packed, const void* w_s, const void* w_zp, void* out,
                 int64_t M, int64_t K, int64_t N,
                 int weight_size_bits, int weight_bias, int64_t group_size,
                 int act_dtype, int out_dtype, int zp_kind, int zp_is_scalar,
                 void* workspace, size_t workspace_bytes,
                 int variant, void* stream, void* const* peer_out, int num_peers);

/* cb200_mpgemm_push with per-GEMM visibility (no cross-GPU barrier): once the WHOLE [M, N] result has been stored to `out`
 * and to every peer, the kernel INCREMENTS local_flag and peer_flags[i] by one (system-scope atomics) -- 32-bit counters,
 * peer_flags[i] a device pointer into peer i's memory (typically slot `rank` of a [world] uint32 array in symmetric memory,
 * zero-initialised).  A consumer on any GPU that needs the gathered output waits until all `world` counters of ITS array have
 * reached the number of signalled GEMVs so far: cb200_wait_flags keeps that number in `*expected` (device memory, zero-
 * initialised, advanced by the wait itself), so a signalled GEMV followed by ONE wait on every rank is safe to capture in a
 * CUDA graph and replay.  Decode kernels only (AUTO: M <= 8, or <= 16 where the streaming kernel supports the problem; M <= 16 with variant = CB200_DECODE_MMA / CB200_DECODE_STREAM).
 * local_flag == NULL: plain cb200_mpgemm_push. */
CB200_API int cb200_mpgemm_push_signal(const void* x, const void* w_q_packed, const void* w_s, const void* w_zp, void* out,
                 int64_t M, int64_t K, int64_t N,
                 int weight_size_bits, int weight_bias, int64_t group_size,
                 int act_dtype, int out_dtype, int zp_kind, int zp_is_scalar,
                 void* workspace, size_t workspace_bytes,
                 int variant, void* stream, void* const* peer_out, int num_peers,
                 void* const* peer_flags, void* local_flag);
/* Enqueue a one-warp kernel on `stream` that advances *expected by one and returns once flags[0 .. count) have all reached it
 * (count <= 32). */
CB200_API int cb200_wait_flags(const void* flags, int count, void* expected, void* stream);

/* cb200_mpgemm_push_signal whose ACTIVATIONS are themselves a gathered buffer that other GPUs push into (the all-gather -> GEMM
 * side of an N-sharded chain: gate_up -> gather -> down -> gather ...).  x is complete once x_flags[0 .. x_flag_count) (the
 * completion counters the producing cb200_mpgemm_push_signal / cb200_push calls raise on this GPU) have all reached
 * *x_expected + 1.  The WAIT HAPPENS INSIDE THE KERNEL: its consumer warps spin on the counters while the producer warp already
 * fills the shared-memory ring with this GEMV's weights, so the NVLink latency of the previous GEMV's push overlaps this one's
 * prologue -- no wait kernel, no barrier, no NCCL.  The grid's last block advances *x_expected (device memory, zero-initialised),
 * so one chained call per signalled producer is safe under CUDA-graph replay.  Streaming decode kernel only (M <= 8; AUTO
 * selects it).  peer_out / peer_flags / local_flag as in cb200_mpgemm_push_signal (num_peers == 0 and local_flag == NULL: the
 * result stays local).  x_flags == NULL: plain cb200_mpgemm_push_signal. */
CB200_API int cb200_mpgemm_chain(const void* x, const void* w_q_packed, const void* w_s, const void* w_zp, void* out,
                 int64_t M, int64_t K, int64_t N,
                 int weight_size_bits, int weight_bias, int64_t group_size,
                 int act_dtype, int out_dtype, int zp_kind, int zp_is_scalar,
                 void* workspace, size_t workspace_bytes,
                 int variant, void* stream, void* const* peer_out, int num_peers,
                 void* const* peer_flags, void* local_flag,
                 const void* x_flags, int x_flag_count, void* x_expected);

/* The gather WITHOUT the GEMM: copy `bytes` (a multiple of 16) from src to peer_dst[0 .. num_peers) with 16-byte P2P stores and
 * -- when local_flag != NULL -- raise the same completion counters as cb200_mpgemm_push_signal once everything has been stored.
 * Meant for a SIDE stream: the kernel uses no shared memory and becomes resident next to the decode kernel of the main stream,
 * so a token's GEMV chain contains no kernel with outstanding NVLink stores (a grid is not complete, and its dependents are
 * not released, before its remote stores are acknowledged).  A signalled push is one CTA (<= 1 MiB). */
CB200_API int cb200_push(const void* src, size_t bytes, void* const* peer_dst, int num_peers,
                 void* const* peer_flags, void* local_flag, void* stream);

/* The large-M leg of the gather: the same copy by up to `max_ctas` CTAs (0 = as many as the bytes are worth, one per SM at most),
 * no completion counters -- close it with a cross-GPU barrier.  multicast_dst != NULL: one `multimem.st` per 16 bytes to an NVLS
 * multicast address that maps this rank's slot of EVERY GPU's buffer (torch symmetric memory: handle.multicast_ptr + slot
 * offset) -- the switch replicates, the sender's egress is the slice once; peer_dst is then ignored.  Meant for a HIGH-PRIORITY
 * side stream next to the following GEMM: the GEMM's CTAs carry 60 K registers, so the push needs SMs of its own and gets them
 * as GEMM CTAs retire (conch_b200/sharded.py gathered_push_async). */
CB200_API int cb200_push_bulk(const void* src, size_t bytes, void* const* peer_dst, int num_peers, void* multicast_dst,
                 int max_ctas, void* stream);

/* cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, stream): the copy-engine leg of the large-M all-gather -- every rank
 * copies its [M, N/G] slice into its slot of every peer's buffer from per-peer side streams while the next GEMM owns the SMs
 * (conch_b200/sharded.py gathered_ce_async). */
CB200_API int cb200_peer_copy(void* dst, const void* src, size_t bytes, void* stream);

/* Which kernel family the last cb200_mpgemm call on this thread dispatched to (a cb200_variant), and how many
 * kernel launches it enqueued.  Used by tests and by bench.py's `gpu_launches` accounting. */
CB200_API int cb200_last_variant(void);
CB200_API int cb200_last_launch_count(void);

/* w_out[K,N] (act dtype) = dequant(w_q_packed, w_s, w_zp): the `_dequantize` stage alone
 * (reference conch/kernels/quantization/gemm.py:176-216) -- bit-exact against the reference's w_ref. */
CB200_API int cb200_dequantize(const void* w_q_packed, const void* w_s, const void* w_zp, void* w_out,
                     int64_t K, int64_t N,
                     int weight_size_bits, int weight_bias, int64_t group_size,
                     int act_dtype, int zp_kind, int zp_is_scalar, void* stream);

/* out[M,N] = cast( (a[M,K] @ b[K,N]) * (scale_a[m] * scale_b[n]) ) (+ bias[n]): W8A8 GEMM with per-token x per-channel scales.
 *   a        [M, K] int8 | fp8 e4m3, row-major
 *   b_t      [N, K] same type, row-major -- the reference's `b` ([K, N]) in the layout its callers hand it, b = w.T with w
 *            contiguous (tests/scaled_gemm_test.py:68-72); the host side transposes anything else once
 *   scale_a  fp32 [M] or 1 element (scale_a_is_scalar); scale_b fp32 [N] or 1 element
 *   bias     [N] in the output dtype or NULL; added to the ROUNDED product and rounded again, as the reference's
 *            `output.add_(bias)` does (conch/ops/quantization/gemm.py:249-250)
 *   out      [M, N] fp16 | bf16 | fp32
 * K % 16 == 0, N % 8 == 0, every pointer 16-byte aligned.  int8 results are bit-identical to the reference op's arithmetic
 * (exact int32 accumulation, acc * (scale_a * scale_b) in fp32, one rounding).
 * Replaces: scaled_gemm_launcher(output, a, b, scale_a, scale_b, metadata) (reference conch/kernels/quantization/gemm.py:564-627). */
CB200_API int cb200_scaled_gemm(const void* a, const void* b_t, const void* scale_a, const void* scale_b, const void* bias, void* out,
                 int64_t M, int64_t K, int64_t N, int in_kind, int out_dtype,
                 int scale_a_is_scalar, int scale_b_is_scalar, void* stream);

/* out[i] = quantise(x[i] * (1 / scale[0])) for a whole tensor of n elements: int8 (clamp to [-128, 127], truncation toward
 * zero) or fp8 e4m3 (clamp to +-448, round to nearest even).  x fp16 | bf16 | fp32 (in_dtype), scale fp32 [1], out 1 byte per
 * element.  x must be 16-byte aligned, out 8-byte aligned.
 * Replaces: static_scaled_int8_quant_launcher / static_scaled_fp8_quant_launcher
 *           (reference conch/kernels/quantization/int8.py:61-97, fp8.py:67-97). */
CB200_API int cb200_static_scaled_quant(const void* x, const void* scale, void* out, int64_t n, int in_dtype, int out_kind, void* stream);

/* Debug/tuning knobs of the decode kernel (0 = heuristic).  Process-global, not thread-safe.
 *   warps_n   1: dequantise every weight before the MMA (W bits equal the reference's w_ref); 0/2: fused operands (default)
 *   warps_k   CTAs per SM assumed resident when the split count is chosen
 *   split_k   forced number of K slabs per 32-column block (rounded to a legal divisor)
 *   vec_words 1: element (2-byte) stores to the peers in cb200_mpgemm_push instead of 8-byte vectors (A/B measurements) */
CB200_API void cb200_set_decode_tuning(int warps_n, int warps_k, int split_k, int vec_words);
/* Debug/tuning knobs for the tcgen05 family: m_tile (MMA N: 16..256, 0 = heuristic), split_k (0 = heuristic). */
CB200_API void cb200_set_tc_tuning(int m_tile, int split_k);
/* CTAs per cluster sharing one activation tile through TMA multicast in the tcgen05 family (1, 2, 4; 0 = default = 1:
 * measured neutral). */
CB200_API void cb200_set_tc_cluster(int ctas);

/* Debug/tuning knobs of the streaming decode kernel (0 = heuristic): shared memory per CTA in KiB (default: all of it),
 * upper bound on the number of CTAs (default: one per SM).  Process-global, not thread-safe. */
CB200_API void cb200_set_stream_tuning(int ring_kb, int ctas);
/* Timeline of the streaming decode kernel: when non-NULL, every following launch stamps %globaltimer at its milestones into
 * device_buffer[cta * 8 + slot] (int64; slots documented in benchmarks/stream_trace.py; debug builds -DCB200_TC_TRACE only).
 * NULL switches it off. */
CB200_API void cb200_debug_set_stream_trace(void* device_buffer);

/* Programmatic dependent launch for the decode kernels (default off).  When enabled, a decode kernel may start while the
 * previous kernel in the stream is still draining and prefetches its WEIGHTS, scales and zero-points before waiting for
 * that kernel; x, out and the workspace are only touched afterwards.  The caller thereby asserts that w_q_packed / w_s /
 * w_zp are not written by the kernel enqueued immediately before the call (true for inference: weights are static). */
CB200_API void cb200_set_pdl(int enable);

#ifdef __cplusplus
}
#endif
#endif /* CONCH_B200_H_ */
