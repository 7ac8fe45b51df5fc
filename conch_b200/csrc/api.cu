// extern "C" entry points of libconch_b200.so (declared in include/conch_b200.h) + regime dispatch.
// Replaces the reference's launcher + dispatch (conch/kernels/quantization/gemm.py:482-545 and the platform
// switch in conch/platforms/__init__.py:14-19): there is exactly one backend, sm_100a.
#include <atomic>
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "host_util.cuh"

namespace cb200 {

static thread_local char tls_error[512] = {0};
static thread_local int tls_last_variant = 0;
static thread_local int tls_last_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tls_error, sizeof(tls_error), fmt, ap);
  va_end(ap);
}

// ---- per-device facts (one process may drive several GPUs: nothing here is cached per process) ---------------------
const DeviceInfo* current_device_info() {
  static DeviceInfo table[CB200_MAX_DEVICES];
  static std::atomic<int> state[CB200_MAX_DEVICES];  // 0 unknown, 1 valid
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= CB200_MAX_DEVICES) return nullptr;
  if (state[dev].load(std::memory_order_acquire) != 1) {
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (state[dev].load(std::memory_order_relaxed) != 1) {
      DeviceInfo di;
      di.index = dev;
      if (cudaDeviceGetAttribute(&di.sm_major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
          cudaDeviceGetAttribute(&di.sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
        return nullptr;
      }
      table[dev] = di;
      state[dev].store(1, std::memory_order_release);
    }
  }
  return &table[dev];
}

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {  // a driver entry point, not a device property: process-wide is correct
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
  });
  return fn;
}

static int check_device() {
  const DeviceInfo* di = current_device_info();
  if (!di || di->sm_major != 10) {
    set_error("libconch_b200 requires an sm_100 (B200) device; there is no fallback path");
    return CB200_ERR_DEVICE;
  }
  return 0;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int validate(const void* x, const void* wq, const void* ws, const void* wzp, const void* out, int64_t M,
                    int64_t K, int64_t N, int bits, int64_t* group_size, int act_dtype, int out_dtype, int zp_kind,
                    bool zp_scalar, bool need_x) {
  if ((need_x && !x) || !wq || !ws || !out) {
    set_error("null pointer argument");
    return CB200_ERR_ARG;
  }
  if ((need_x && M <= 0) || K <= 0 || N <= 0) {
    set_error("non-positive dimension M=%lld K=%lld N=%lld", (long long)M, (long long)K, (long long)N);
    return CB200_ERR_ARG;
  }
  if (bits != 4 && bits != 8) {
    set_error("weight_size_bits=%d not supported (4 or 8)", bits);
    return CB200_ERR_UNSUPPORTED;
  }
  if (act_dtype != CB200_F16 && act_dtype != CB200_BF16) {
    set_error("activation dtype %d not supported (fp16 or bf16)", act_dtype);
    return CB200_ERR_UNSUPPORTED;
  }
  if (out_dtype != CB200_F16 && out_dtype != CB200_BF16 && out_dtype != CB200_F32) {
    set_error("output dtype %d not supported", out_dtype);
    return CB200_ERR_UNSUPPORTED;
  }
  if (zp_kind < CB200_ZP_NONE || zp_kind > CB200_ZP_ACT || (zp_kind != CB200_ZP_NONE && !wzp)) {
    set_error("bad zero-point kind %d / pointer", zp_kind);
    return CB200_ERR_ARG;
  }
  if (*group_size == -1) *group_size = K;
  const int64_t g = *group_size;
  if (K % 64 != 0 || N % 64 != 0) {
    set_error("K=%lld and N=%lld must be multiples of 64 (the reference kernel's unmasked 64x64 weight tile)",
              (long long)K, (long long)N);
    return CB200_ERR_SHAPE;
  }
  if (g <= 0 || g > K || K % g != 0 || (g % 64 != 0)) {
    set_error("group_size=%lld must divide K=%lld and be a multiple of 64 (or -1 / K for per-channel)",
              (long long)g, (long long)K);
    return CB200_ERR_SHAPE;
  }
  if (M > (1 << 24) || K > (1 << 24) || N > (1 << 24)) {
    set_error("dimension too large");
    return CB200_ERR_SHAPE;
  }
  // (per-group zero-points are read with 16-byte loads like the scales; a scalar zero-point may sit anywhere)
  if ((need_x && !aligned16(x)) || !aligned16(wq) || !aligned16(out) || !aligned16(ws) ||
      (wzp && zp_kind != CB200_ZP_NONE && !zp_scalar && !aligned16(wzp))) {
    set_error("x, w_q_packed, w_s, per-group w_zp and out must be 16-byte aligned");
    return CB200_ERR_ARG;
  }
  return 0;
}

}  // namespace cb200

using namespace cb200;

extern "C" {

int cb200_version(void) { return CB200_VERSION; }
const char* cb200_last_error(void) { return tls_error; }
int cb200_last_variant(void) { return tls_last_variant; }
int cb200_last_launch_count(void) { return tls_last_launches; }

void cb200_set_decode_tuning(int warps_n, int warps_k, int split_k, int vec_words) {
  g_decode_tuning = {warps_n, warps_k, split_k, vec_words};
}
void cb200_set_tc_tuning(int m_tile, int split_k) { g_tc_tuning = {m_tile, split_k}; }
void cb200_set_tc_cluster(int ctas) { g_tc_cluster = (ctas == 1 || ctas == 2 || ctas == 4) ? ctas : 0; }
void cb200_set_pdl(int enable) { g_decode_pdl = enable ? 1 : 0; }
void cb200_set_stream_tuning(int ring_kb, int ctas) {
  // (ring_kb < 0: debug -- |ring_kb| with the weight tiles loaded WITHOUT the L2 evict_first policy, for A/B timing)
  g_stream_tuning = {ring_kb < 0 ? -ring_kb : ring_kb, ctas, ring_kb < 0 ? 1 : 0};
}
void cb200_debug_set_stream_trace(void* device_buffer) { g_stream_trace = reinterpret_cast<long long*>(device_buffer); }

size_t cb200_mpgemm_workspace_bytes(int64_t M, int64_t K, int64_t N, int weight_size_bits, int variant) {
  (void)variant;
  const int bits = weight_size_bits == 8 ? 8 : 4;
  const size_t a = decode_workspace_bytes(M, K, N, bits);
  const size_t b = tc_workspace_bytes(M, K, N);
  const size_t c = stream_workspace_bytes(M, K, N, bits);
  size_t need = a > b ? a : b;
  return c > need ? c : need;
}

int cb200_dequantize(const void* w_q_packed, const void* w_s, const void* w_zp, void* w_out, int64_t K, int64_t N,
                     int weight_size_bits, int weight_bias, int64_t group_size, int act_dtype, int zp_kind,
                     int zp_is_scalar, void* stream) {
  int rc = check_device();
  if (rc) return rc;
  rc = validate(nullptr, w_q_packed, w_s, w_zp, w_out, 1, K, N, weight_size_bits, &group_size, act_dtype, act_dtype,
                zp_kind, zp_is_scalar != 0, false);
  if (rc) return rc;
  Problem p{};
  p.wq = reinterpret_cast<const uint32_t*>(w_q_packed);
  p.s = w_s;
  p.zp = w_zp;
  p.out = w_out;
  p.M = 0;
  p.K = int(K);
  p.N = int(N);
  p.bias = weight_bias;
  p.g = int(group_size);
  p.zp_kind = zp_kind;
  p.zp_scalar = zp_is_scalar;
  p.out_dtype = act_dtype;
  return launch_dequant(p, weight_size_bits, act_dtype, reinterpret_cast<cudaStream_t>(stream));
}

int cb200_mpgemm(const void* x, const void* w_q_packed, const void* w_s, const void* w_zp, void* out, int64_t M,
                 int64_t K, int64_t N, int weight_size_bits, int weight_bias, int64_t group_size, int act_dtype,
                 int out_dtype, int zp_kind, int zp_is_scalar, void* workspace, size_t workspace_bytes, int variant,
                 void* stream) {
  return cb200_mpgemm_push(x, w_q_packed, w_s, w_zp, out, M, K, N, weight_size_bits, weight_bias, group_size, act_dtype,
                           out_dtype, zp_kind, zp_is_scalar, workspace, workspace_bytes, variant, stream, nullptr, 0);
}

int cb200_mpgemm_push(const void* x, const void* w_q_packed, const void* w_s, const void* w_zp, void* out, int64_t M,
                      int64_t K, int64_t N, int weight_size_bits, int weight_bias, int64_t group_size, int act_dtype,
                      int out_dtype, int zp_kind, int zp_is_scalar, void* workspace, size_t workspace_bytes, int variant,
                      void* stream, void* const* peer_out, int num_peers) {
  return cb200_mpgemm_push_signal(x, w_q_packed, w_s, w_zp, out, M, K, N, weight_size_bits, weight_bias, group_size, act_dtype,
                                  out_dtype, zp_kind, zp_is_scalar, workspace, workspace_bytes, variant, stream, peer_out, num_peers,
                                  nullptr, nullptr);
}

int cb200_wait_flags(const void* flags, int count, void* expected, void* stream) {
  if (!flags || !expected || count < 1 || count > 32) {
    set_error("cb200_wait_flags: null pointer or count=%d outside 1..32", count);
    return CB200_ERR_ARG;
  }
  return launch_wait_flags(reinterpret_cast<const uint32_t*>(flags), count, reinterpret_cast<uint32_t*>(expected),
                           reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"

static int run_mpgemm(const void* x, const void* w_q_packed, const void* w_s, const void* w_zp, void* out, int64_t M, int64_t K,
                      int64_t N, int weight_size_bits, int weight_bias, int64_t group_size, int act_dtype, int out_dtype,
                      int zp_kind, int zp_is_scalar, void* workspace, size_t workspace_bytes, int variant, void* stream,
                      void* const* peer_out, int num_peers, void* const* peer_flags, void* local_flag, const void* bias,
                      const void* x_flags = nullptr, int x_flag_count = 0, void* x_expected = nullptr) {
  tls_last_variant = 0;
  tls_last_launches = 0;
  int rc = check_device();
  if (rc) return rc;
  rc = validate(x, w_q_packed, w_s, w_zp, out, M, K, N, weight_size_bits, &group_size, act_dtype, out_dtype, zp_kind,
                zp_is_scalar != 0, true);
  if (rc) return rc;
  if (workspace && (reinterpret_cast<uintptr_t>(workspace) & 15u)) {
    set_error("workspace must be 16-byte aligned");
    return CB200_ERR_ARG;
  }
  Problem p{};
  p.x = x;
  p.wq = reinterpret_cast<const uint32_t*>(w_q_packed);
  p.s = w_s;
  p.zp = w_zp;
  p.out = out;
  p.M = int(M);
  p.K = int(K);
  p.N = int(N);
  p.bias = weight_bias;
  p.g = int(group_size);
  p.zp_kind = zp_kind;
  p.zp_scalar = zp_is_scalar;
  p.out_dtype = out_dtype;
  p.out_bias = bias;
  if (bias != nullptr && (reinterpret_cast<uintptr_t>(bias) & (out_dtype == CB200_F32 ? 3u : 1u))) {
    set_error("bias is not aligned to its element size");
    return CB200_ERR_ARG;
  }
  if (num_peers < 0 || num_peers > CB200_MAX_PEERS || (num_peers > 0 && !peer_out)) {
    set_error("num_peers=%d out of range (0..%d)", num_peers, CB200_MAX_PEERS);
    return CB200_ERR_ARG;
  }
  p.n_peers = num_peers;
  for (int i = 0; i < num_peers; ++i) {
    if (!peer_out[i] || !aligned16(peer_out[i])) {
      set_error("peer_out[%d] is null or not 16-byte aligned", i);
      return CB200_ERR_ARG;
    }
    p.peer_out[i] = peer_out[i];
  }
  if (local_flag != nullptr) {
    if (num_peers > 0 && !peer_flags) {
      set_error("cb200_mpgemm_push_signal: peer_flags is null");
      return CB200_ERR_ARG;
    }
    for (int i = 0; i < num_peers; ++i) {
      if (!peer_flags[i] || (reinterpret_cast<uintptr_t>(peer_flags[i]) & 3u)) {
        set_error("peer_flags[%d] is null or not 4-byte aligned", i);
        return CB200_ERR_ARG;
      }
      p.peer_flag[i] = reinterpret_cast<uint32_t*>(peer_flags[i]);
    }
    p.local_flag = reinterpret_cast<uint32_t*>(local_flag);
  }
  if (x_flags != nullptr) {
    if (!x_expected || x_flag_count < 1 || x_flag_count > 32 || (reinterpret_cast<uintptr_t>(x_flags) & 3u) ||
        (reinterpret_cast<uintptr_t>(x_expected) & 3u)) {
      set_error("cb200_mpgemm_chain: x_expected is null / misaligned or x_flag_count=%d outside 1..32", x_flag_count);
      return CB200_ERR_ARG;
    }
    p.wait_flags = reinterpret_cast<const uint32_t*>(x_flags);
    p.wait_expected = reinterpret_cast<uint32_t*>(x_expected);
    p.wait_count = x_flag_count;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  int chosen = variant;
  if (chosen == CB200_AUTO) {
    // regime split by MEASUREMENT (profiles/r02_*): HBM-bound skinny GEMM up to 8 rows (one 8-wide MMA tile of x) on the
    // streaming ring kernel where it has an operand format and a stage layout for the problem, else on round 1's
    // register-window kernel (group sizes that straddle stages); tcgen05 pipeline above.
    // (M = 9..16: the streaming kernel's 16-row variant where it has a format for the problem -- 3 to 17 % ahead of the tcgen05
    //  kernel on the four Llama-3-8B shapes, profiles/r02_decode_table_m16.log; otherwise tcgen05, as for everything above)
    //  8-bit weights keep M = 9..16 on tcgen05: twice the MMAs per weight, 56 against 48 us on gate_up, profiles/r02_config4_*.log)
    const bool decode = M <= 8 || (M <= 16 && weight_size_bits == 4 && stream_supported(p, weight_size_bits, act_dtype));
    if (decode || !tc_supported(p, weight_size_bits)) {
      // (fused gather: the streaming kernel for any number of peers -- 8 GPUs, 128 pushing GEMVs per token: 1.35 ms against 1.84 ms
      //  on the register-window kernel and 1.31 ms without any push, profiles/r02_push_modes_n8.log)
      // (the in-kernel wait of cb200_mpgemm_chain exists in the streaming kernel only)
      chosen = (decode && stream_supported(p, weight_size_bits, act_dtype)) ? CB200_DECODE_STREAM : CB200_DECODE_MMA;
    } else {
      chosen = CB200_TC_TS;
    }
  }
  if (num_peers > 0 && chosen != CB200_DECODE_MMA && chosen != CB200_DECODE_STREAM) {
    // the fused push lives in the decode kernels' (tiny) epilogue; big outputs are bandwidth-bound: all-gather them
    set_error("cb200_mpgemm_push: only the decode kernels (M <= 16) store to peers; use an all-gather for M=%lld", (long long)M);
    return CB200_ERR_UNSUPPORTED;
  }
  if (p.local_flag != nullptr && ((chosen != CB200_DECODE_MMA && chosen != CB200_DECODE_STREAM) || M > 16)) {
    set_error("cb200_mpgemm_push_signal: the completion flags are raised by the decode kernels only (M <= 16 in one pass)");
    return CB200_ERR_UNSUPPORTED;
  }
  if (p.wait_flags != nullptr && chosen != CB200_DECODE_STREAM) {
    set_error("cb200_mpgemm_chain: only the streaming decode kernel waits for its activations (M <= 8, fp16/bf16 x int4 or fp16 x int8)");
    return CB200_ERR_UNSUPPORTED;
  }
  int launches = 0;
  if (chosen == CB200_DECODE_MMA) {
    rc = launch_decode(p, weight_size_bits, act_dtype, workspace, workspace_bytes, st, &launches);
  } else if (chosen == CB200_DECODE_STREAM) {
    if (!stream_supported(p, weight_size_bits, act_dtype)) {
      set_error("the streaming decode kernel does not support this problem (a group size that is neither 1, 2 nor a multiple of 4 quads of packed rows: use CB200_DECODE_MMA)");
      return CB200_ERR_UNSUPPORTED;
    }
    rc = launch_stream(p, weight_size_bits, act_dtype, workspace, workspace_bytes, st, &launches);
  } else if (chosen == CB200_TC_SS || chosen == CB200_TC_TS || chosen == CB200_TC2_TS) {
    if (!tc_supported(p, weight_size_bits)) {
      set_error("tcgen05 variant %d does not support this problem", chosen);
      return CB200_ERR_UNSUPPORTED;
    }
    if (chosen == CB200_TC2_TS) {
      rc = launch_tc2(p, weight_size_bits, act_dtype, st, &launches);  // persistent, m_tile 128 (large M)
    } else {
      rc = launch_tc(p, weight_size_bits, act_dtype, chosen, workspace, workspace_bytes, st, &launches);
    }
  } else {
    set_error("unknown variant %d", variant);
    return CB200_ERR_ARG;
  }
  tls_last_variant = chosen;
  tls_last_launches = launches;
  return rc;
}

extern "C" {

int cb200_mpgemm_push_signal(const void* x, const void* w_q_packed, const void* w_s, const void* w_zp, void* out, int64_t M,
                             int64_t K, int64_t N, int weight_size_bits, int weight_bias, int64_t group_size, int act_dtype,
                             int out_dtype, int zp_kind, int zp_is_scalar, void* workspace, size_t workspace_bytes, int variant,
                             void* stream, void* const* peer_out, int num_peers, void* const* peer_flags, void* local_flag) {
  return run_mpgemm(x, w_q_packed, w_s, w_zp, out, M, K, N, weight_size_bits, weight_bias, group_size, act_dtype, out_dtype, zp_kind,
                    zp_is_scalar, workspace, workspace_bytes, variant, stream, peer_out, num_peers, peer_flags, local_flag, nullptr);
}

int cb200_mpgemm_chain(const void* x, const void* w_q_packed, const void* w_s, const void* w_zp, void* out, int64_t M, int64_t K,
                       int64_t N, int weight_size_bits, int weight_bias, int64_t group_size, int act_dtype, int out_dtype, int zp_kind,
                       int zp_is_scalar, void* workspace, size_t workspace_bytes, int variant, void* stream, void* const* peer_out,
                       int num_peers, void* const* peer_flags, void* local_flag, const void* x_flags, int x_flag_count,
                       void* x_expected) {
  return run_mpgemm(x, w_q_packed, w_s, w_zp, out, M, K, N, weight_size_bits, weight_bias, group_size, act_dtype, out_dtype, zp_kind,
                    zp_is_scalar, workspace, workspace_bytes, variant, stream, peer_out, num_peers, peer_flags, local_flag, nullptr,
                    x_flags, x_flag_count, x_expected);
}

int cb200_push(const void* src, size_t bytes, void* const* peer_dst, int num_peers, void* const* peer_flags, void* local_flag,
               void* stream) {
  int rc = check_device();
  if (rc) return rc;
  if (!src || !aligned16(src) || bytes == 0 || (bytes & 15u)) {
    set_error("cb200_push: src must be 16-byte aligned and bytes a positive multiple of 16 (bytes=%zu)", bytes);
    return CB200_ERR_ARG;
  }
  if (num_peers < 1 || num_peers > CB200_MAX_PEERS || !peer_dst || (local_flag && !peer_flags)) {
    set_error("cb200_push: num_peers=%d out of range (1..%d) or null pointer array", num_peers, CB200_MAX_PEERS);
    return CB200_ERR_ARG;
  }
  for (int i = 0; i < num_peers; ++i) {
    if (!peer_dst[i] || !aligned16(peer_dst[i]) || (local_flag && (!peer_flags[i] || (reinterpret_cast<uintptr_t>(peer_flags[i]) & 3u)))) {
      set_error("cb200_push: peer_dst[%d] null / not 16-byte aligned, or peer_flags[%d] null / misaligned", i, i);
      return CB200_ERR_ARG;
    }
  }
  if (local_flag && bytes > (size_t(1) << 20)) {
    set_error("cb200_push: a signalled push is one CTA (decode-sized slices, <= 1 MiB); push %zu bytes without flags", bytes);
    return CB200_ERR_UNSUPPORTED;
  }
  return launch_push(src, bytes, peer_dst, num_peers, peer_flags, local_flag, reinterpret_cast<cudaStream_t>(stream));
}

int cb200_push_bulk(const void* src, size_t bytes, void* const* peer_dst, int num_peers, void* multicast_dst, int max_ctas,
                    void* stream) {
  int rc = check_device();
  if (rc) return rc;
  if (!src || !aligned16(src) || bytes == 0 || (bytes & 15u)) {
    set_error("cb200_push_bulk: src must be 16-byte aligned and bytes a positive multiple of 16 (bytes=%zu)", bytes);
    return CB200_ERR_ARG;
  }
  if (multicast_dst != nullptr) {
    if (!aligned16(multicast_dst)) {
      set_error("cb200_push_bulk: multicast_dst is not 16-byte aligned");
      return CB200_ERR_ARG;
    }
  } else {
    if (num_peers < 1 || num_peers > CB200_MAX_PEERS || !peer_dst) {
      set_error("cb200_push_bulk: num_peers=%d out of range (1..%d) or null pointer array", num_peers, CB200_MAX_PEERS);
      return CB200_ERR_ARG;
    }
    for (int i = 0; i < num_peers; ++i) {
      if (!peer_dst[i] || !aligned16(peer_dst[i])) {
        set_error("cb200_push_bulk: peer_dst[%d] null or not 16-byte aligned", i);
        return CB200_ERR_ARG;
      }
    }
  }
  return launch_push_bulk(src, bytes, peer_dst, num_peers, multicast_dst, max_ctas, reinterpret_cast<cudaStream_t>(stream));
}

int cb200_peer_copy(void* dst, const void* src, size_t bytes, void* stream) {
  if (!dst || !src) {
    set_error("cb200_peer_copy: null pointer");
    return CB200_ERR_ARG;
  }
  if (bytes == 0) return 0;
  CB200_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int cb200_mpgemm_bias(const void* x, const void* w_q_packed, const void* w_s, const void* w_zp, const void* bias, void* out,
                      int64_t M, int64_t K, int64_t N, int weight_size_bits, int weight_bias, int64_t group_size, int act_dtype,
                      int out_dtype, int zp_kind, int zp_is_scalar, void* workspace, size_t workspace_bytes, int variant,
                      void* stream) {
  return run_mpgemm(x, w_q_packed, w_s, w_zp, out, M, K, N, weight_size_bits, weight_bias, group_size, act_dtype, out_dtype, zp_kind,
                    zp_is_scalar, workspace, workspace_bytes, variant, stream, nullptr, 0, nullptr, nullptr, bias);
}

int cb200_scaled_gemm(const void* a, const void* b_t, const void* scale_a, const void* scale_b, const void* bias, void* out, int64_t M,
                      int64_t K, int64_t N, int in_kind, int out_dtype, int scale_a_is_scalar, int scale_b_is_scalar, void* stream) {
  int rc = check_device();
  if (rc) return rc;
  if (!a || !b_t || !scale_a || !scale_b || !out) {
    set_error("null pointer argument");
    return CB200_ERR_ARG;
  }
  if (M <= 0 || K <= 0 || N <= 0 || M > (1 << 24) || K > (1 << 24) || N > (1 << 24)) {
    set_error("dimension out of range M=%lld K=%lld N=%lld", (long long)M, (long long)K, (long long)N);
    return CB200_ERR_ARG;
  }
  if (in_kind != CB200_IN_INT8 && in_kind != CB200_IN_FP8_E4M3) {
    set_error("scaled_gemm input kind %d not supported (int8 or fp8 e4m3)", in_kind);
    return CB200_ERR_UNSUPPORTED;
  }
  if (out_dtype != CB200_F16 && out_dtype != CB200_BF16 && out_dtype != CB200_F32) {
    set_error("output dtype %d not supported", out_dtype);
    return CB200_ERR_UNSUPPORTED;
  }
  if (K % 16 != 0 || N % 8 != 0) {
    set_error("scaled_gemm needs K %% 16 == 0 and N %% 8 == 0 (K=%lld, N=%lld): rows are fetched by TMA / stored as 16-byte vectors",
              (long long)K, (long long)N);
    return CB200_ERR_SHAPE;
  }
  if (!aligned16(a) || !aligned16(b_t) || !aligned16(out) || (reinterpret_cast<uintptr_t>(scale_a) & 3u) ||
      (reinterpret_cast<uintptr_t>(scale_b) & 3u)) {
    set_error("a, b and out must be 16-byte aligned, the scales 4-byte aligned");
    return CB200_ERR_ARG;
  }
  return launch_scaled(a, b_t, reinterpret_cast<const float*>(scale_a), reinterpret_cast<const float*>(scale_b), bias, out, int(M), int(K),
                       int(N), in_kind, out_dtype, scale_a_is_scalar, scale_b_is_scalar, reinterpret_cast<cudaStream_t>(stream));
}

int cb200_static_scaled_quant(const void* x, const void* scale, void* out, int64_t n, int in_dtype, int out_kind, void* stream) {
  int rc = check_device();
  if (rc) return rc;
  if (!x || !scale || !out || n < 0) {
    set_error("null pointer / negative size");
    return CB200_ERR_ARG;
  }
  if (in_dtype != CB200_F16 && in_dtype != CB200_BF16 && in_dtype != CB200_F32) {
    set_error("input dtype %d not supported (fp16, bf16, fp32)", in_dtype);
    return CB200_ERR_UNSUPPORTED;
  }
  if (out_kind != CB200_IN_INT8 && out_kind != CB200_IN_FP8_E4M3) {
    set_error("output kind %d not supported (int8 or fp8 e4m3)", out_kind);
    return CB200_ERR_UNSUPPORTED;
  }
  if (!aligned16(x) || (reinterpret_cast<uintptr_t>(out) & 7u) || (reinterpret_cast<uintptr_t>(scale) & 3u)) {
    set_error("x must be 16-byte aligned, out 8-byte aligned, scale 4-byte aligned");
    return CB200_ERR_ARG;
  }
  if (n == 0) return 0;
  return launch_static_quant(x, reinterpret_cast<const float*>(scale), out, n, in_dtype, out_kind, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
