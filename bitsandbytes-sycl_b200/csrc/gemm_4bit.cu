// gemm_4bit.cu -- K4: fused batch>1 4-bit GEMM, out[b, n] = sum_k A[b,k] * T(code[q(n,k)] * absmax[(n*K+k)/bs]) (+ bias[n]).
//
// Replaces the reference's batch>1 route MatMul4Bit.forward (python_src_quants/autograd/_functions.py:490-518):
// dequantize_4bit (kDequantizeBlockwise, kernel_quant.cpp:1370-1471) writes the whole weight to HBM in T, F.linear
// reads it back (2 + 2 bytes per weight on top of the 0.5 that are needed).  Here the packed weight is read once,
// dequantised in registers with EXACTLY the reference's arithmetic -- w = T(fp32 code[q] * fp32 absmax), one
// rounding (kernel_quant.cpp:1449-1450) -- and fed to the 5th-generation tensor cores:
//
//   swap-AB: the 128 x 64 weight tile is the UMMA M-operand (A, K-major, SWIZZLE_128B rows of 128 B written by
//   the dequant warps), the activations [batch, 64] are the N-operand (B, TMA, zero-filled beyond `batch`),
//   D[128 weight rows, batch] accumulates in TMEM (fp32), tcgen05.mma kind::f16, M=128, N=round16(batch), K=16.
//
//   warp 0      TMA producer: the PACKED weight tile (128 rows x 32 B per stage) into a deep ring -- this ring, not the
//               register file, holds the bytes in flight towards HBM (up to 64 KB per CTA) -- and the activation
//               tile (one 128-byte-swizzled box per stage) into the operand ring
//   warp 1      TMEM owner + single-thread tcgen05.mma issuer; tcgen05.commit frees an operand stage
//   warps 2-17  dequant producers (four groups of four warps, group g takes stages kb % 4 == g): thread r owns weight row r of the tile -- two LDS.128 of packed bytes per stage,
//               16-entry fp32 code table in shared memory (16 words in 16 banks: conflict-free for any data),
//               64 FMUL, cvt.rn.{bf16x2,f16x2}.f32, eight swizzled STS.128, fence.proxy.async, mbarrier arrive;
//               after the main loop the same warps run the epilogue (tcgen05.ld -> +bias -> T -> global, or fp32
//               partials for split-K)
//
// A CTA owns one (128-row tile, K split).  Llama-3-8B MLP shapes have only 32..112 row tiles, so K is split until
// the grid fills the SMs (two CTAs per SM for batch <= 64); split partials go to an fp32 workspace in a fixed
// layout and a small kernel sums them in order (deterministic) and applies bias + rounding.
#include <stdio.h>
#include <stdlib.h>
#include <type_traits>

#include <map>
#include <mutex>

#include "common.cuh"
#include "tcgen05.cuh"

namespace bnb {

namespace g4 {
constexpr int TM = 128;            // weight rows per tile (UMMA M)
constexpr int TK = 64;             // K elements per stage (128 bytes of T: one SWIZZLE_128B row)
constexpr int kDqWarps = 16;        // dequant warps: groups of 4 (one thread per weight row), group g takes stages kb % G == g
constexpr int kThreads = 64 + kDqWarps * 32;
constexpr int kStageA = TM * 128;  // 16 KB
constexpr int kMaxNB = 256;
constexpr int kStageW = TM * 32;   // 4 KB of packed weights per stage
constexpr int kMaxWSlots = 16;

struct Args {
  int batch, N, K, blocksize, bs_shift;
  int NB;            // UMMA N: batch rounded up to 16
  int splits, kper;  // K elements per split (multiple of 64)
  int stages;        // operand ring (dequantised A tile + activation tile)
  int wslots;        // packed-weight ring
  const unsigned char *B;
  const float *absmax;
  const float *code;
  const void *bias;  // T[N] or null
  void *out;         // T[batch, N]           (splits == 1)
  float *ws;         // fp32 [splits, batch, N] (splits > 1)
};

template <typename T> __device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&h);
}

__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 2) k_gemm4_tcgen05(const __grid_constant__ CUtensorMap tmX,
                                                            const __grid_constant__ CUtensorMap tmW, const Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stageB = a.NB * 128;
  const int stage_bytes = kStageA + stageB;
  uint8_t *wring = smem + a.stages * stage_bytes;                          // [wslots][128 rows][32 B]
  uint64_t *bars = reinterpret_cast<uint64_t *>(wring + a.wslots * kStageW);
  uint64_t *fullA = bars, *fullB = bars + 8, *empty = bars + 16, *tfull = bars + 24;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 25);
  __shared__ float s_code[16];   // static: the compiler must KNOW these lookups are shared-memory loads (LDS, not generic LD)
  uint64_t *fullW = bars + 34, *emptyW = bars + 34 + kMaxWSlots;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, split = blockIdx.y;
  const int n0 = tile * TM;
  const int k_begin = split * a.kper;
  const int k_end = min(a.K, k_begin + a.kper);
  const int nk = (k_end - k_begin + TK - 1) / TK;     // stages of work (>= 1 by construction)
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < a.NB) tmem_cols <<= 1;

  if (threadIdx.x < 16) s_code[threadIdx.x] = a.code[threadIdx.x];
  if (warp == 0 && lane == 0) { tc::prefetch_tmap(&tmX); tc::prefetch_tmap(&tmW); }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < a.stages; s++) {
        tc::mbar_init(tc::smem_u32(fullA + s), 4);     // one arrival per warp of the group that fills the stage
        tc::mbar_init(tc::smem_u32(fullB + s), 1);     // TMA producer's expect_tx arrival
        tc::mbar_init(tc::smem_u32(empty + s), 1);     // tcgen05.commit
      }
      for (int s = 0; s < a.wslots; s++) {
        tc::mbar_init(tc::smem_u32(fullW + s), 1);     // TMA producer's expect_tx arrival
        tc::mbar_init(tc::smem_u32(emptyW + s), 4);    // one arrival per warp of the group that drains the slot
      }
      tc::mbar_init(tc::smem_u32(tfull), 1);
      tc::fence_barrier_init();
    }
    __syncwarp();
    tc::tmem_alloc(tc::smem_u32(tmem_slot), tmem_cols);
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer: packed weights (deep ring) + activations (operand ring) =================
    if (lane == 0) {
      int stage = 0, wslot = 0; uint32_t phase = 0, wphase = 0;
      const int lead = a.wslots - a.stages;              // the packed ring runs this many stages ahead of the operand ring
      for (int i = 0; i < nk + lead; i++) {
        if (i < nk) {
          tc::mbar_wait(tc::smem_u32(emptyW + wslot), wphase ^ 1);
          const uint32_t fw = tc::smem_u32(fullW + wslot);
          tc::mbar_arrive_expect_tx(fw, kStageW);
          tc::tma_load_2d(tc::smem_u32(wring + wslot * kStageW), &tmW, fw, (k_begin + i * TK) >> 1, n0);
          if (++wslot == a.wslots) { wslot = 0; wphase ^= 1; }
        }
        const int kb = i - lead;
        if (kb >= 0) {
          tc::mbar_wait(tc::smem_u32(empty + stage), phase ^ 1);
          const uint32_t fb = tc::smem_u32(fullB + stage);
          tc::mbar_arrive_expect_tx(fb, (uint32_t)stageB);
          tc::tma_load_2d(tc::smem_u32(smem + stage * stage_bytes + kStageA), &tmX, fb, k_begin + kb * TK, 0);
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = tc::umma_idesc(tc::kCFormatF32, sizeof(T) == 2 && std::is_same<T, __nv_bfloat16>::value ? 1u : 0u, TM, (uint32_t)a.NB);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < nk; kb++) {
        tc::mbar_wait(tc::smem_u32(fullA + stage), phase);
        tc::mbar_wait(tc::smem_u32(fullB + stage), phase);
        tc::fence_after_sync();
        const uint32_t sa = tc::smem_u32(smem + stage * stage_bytes);
        const uint64_t adesc = tc::umma_desc_sw128_kmajor(sa);
        const uint64_t bdesc = tc::umma_desc_sw128_kmajor(sa + kStageA);
#pragma unroll
        for (int k = 0; k < TK / 16; k++)
          tc::umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
        tc::umma_commit(tc::smem_u32(empty + stage));
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
      tc::umma_commit(tc::smem_u32(tfull));
    }
  } else {
    // ================= dequant producers (warps 2..9), then epilogue =================
    // Two groups of 4 warps; a group dequantises every other stage (thread r of the group owns weight row r), so
    // the barrier / fence latencies of one stage overlap the arithmetic of the next.
    constexpr int G = kDqWarps / 4;
    const int dt = threadIdx.x - 64;
    const int r = dt & 127;                               // weight row inside the tile
    const int grp = dt >> 7;
    const int row = min(n0 + r, a.N - 1);                 // clamped: rows past N are computed and dropped
    const size_t ebase = (size_t)row * a.K;
    const uint32_t swz = (uint32_t)(r & 7);
    const uint32_t smem_s = tc::smem_u32(smem), wring_s = tc::smem_u32(wring);
    float am_next = grp < nk ? __ldg(a.absmax + ((ebase + k_begin + grp * TK) >> a.bs_shift)) : 0.f;
    int stage = grp % a.stages, wslot = grp % a.wslots;   // ring positions advance by G per iteration: no divisions in the loop
    uint32_t phase = (uint32_t)(grp / a.stages) & 1u, wphase = (uint32_t)(grp / a.wslots) & 1u;
    for (int kb = grp; kb < nk; kb += G) {
      const float am = am_next;
      if (kb + G < nk) am_next = __ldg(a.absmax + ((ebase + k_begin + (kb + G) * TK) >> a.bs_shift));
      // packed bytes of this row: 32 B out of the TMA-filled ring, then hand the slot back
      tc::mbar_wait(tc::smem_u32(fullW + wslot), wphase);
      const uint32_t wp = wring_s + wslot * kStageW + r * 32;
      const uint4 w0 = lds128(wp), w1 = lds128(wp + 16);
      const uint32_t w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      // the slot is handed back after the stage has been written (every dequantised value consumes the loaded
      // registers, so both loads have returned by then) -- see the note in gemm_4bit_small.cuh
      const uint32_t wbar = tc::smem_u32(emptyW + wslot);
      wslot += G;
      while (wslot >= a.wslots) { wslot -= a.wslots; wphase ^= 1u; }

      tc::mbar_wait(tc::smem_u32(empty + stage), phase ^ 1);
      const uint32_t dst = smem_s + stage * stage_bytes + r * 128;
#pragma unroll
      for (int c = 0; c < 8; c++) {                       // one packed word = 8 elements = one 16-byte chunk
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const uint32_t byte = (w[c] >> (8 * i)) & 0xFFu;
          const float f0 = __fmul_rn(s_code[byte >> 4], am);     // even element: high nibble
          const float f1 = __fmul_rn(s_code[byte & 15u], am);
          o[i] = pack2<T>(f0, f1);
        }
        sts128(dst + (((uint32_t)c ^ swz) << 4), o[0], o[1], o[2], o[3]);
      }
      tc::fence_proxy_async();                            // generic-proxy stores -> visible to the UMMA (async proxy)
      __syncwarp();
      if (lane == 0) {
        tc::mbar_arrive(wbar);
        tc::mbar_arrive(tc::smem_u32(fullA + stage));
      }
      stage += G;
      while (stage >= a.stages) { stage -= a.stages; phase ^= 1u; }
    }

    // ---- epilogue: TMEM lane quarter warp % 4 (two warps per quarter, alternating 32-column groups)
    const int q = warp & 3;
    const int orow = n0 + q * 32 + lane;
    tc::mbar_wait(tc::smem_u32(tfull), 0);
    tc::fence_after_sync();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float bias = 0.f;
    if (a.bias != nullptr && a.splits == 1 && orow < a.N) bias = to_float<T>(reinterpret_cast<const T *>(a.bias)[orow]);
    for (int c0 = ((warp - 2) >> 2) * 32; c0 < a.NB; c0 += 32 * (kDqWarps / 4)) {
      uint32_t v[32];
      tc::tmem_ld_32x32b_x32(taddr + c0, v);              // columns beyond NB are never stored
      tc::tmem_ld_wait();
      if (orow < a.N) {
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const int b = c0 + j;
          if (b < a.batch) {
            const float acc = __uint_as_float(v[j]);
            if (a.splits == 1) reinterpret_cast<T *>(a.out)[(size_t)b * a.N + orow] = from_float<T>(__fadd_rn(acc, bias));
            else a.ws[((size_t)split * a.batch + b) * a.N + orow] = acc;
          }
        }
      }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, tmem_cols);
}

// split-K: out[b, n] = T(sum_s ws[s, b, n] + bias[n]), summed in split order (deterministic)
template <typename T>
__global__ void __launch_bounds__(256) k_gemm4_finalize(const float *__restrict__ ws, const T *__restrict__ bias, T *__restrict__ out,
                                                        int splits, int batch, int N) {
  const size_t total = (size_t)batch * N;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    float s = 0.f;
    for (int k = 0; k < splits; k++) s = __fadd_rn(s, ws[(size_t)k * total + i]);
    if (bias != nullptr) s = __fadd_rn(s, to_float<T>(bias[i % N]));
    out[i] = from_float<T>(s);
  }
}

// split-K partial sums: one workspace per (device, stream) -- launches on one stream are ordered, two streams running
// the kernel concurrently must not share it.  cudaFree of an outgrown buffer synchronises the device first.
struct WsKey { int dev; cudaStream_t st; bool operator<(const WsKey &o) const { return dev != o.dev ? dev < o.dev : st < o.st; } };
struct WsBuf { float *p; size_t bytes; };
static std::mutex g_ws_mu;
static std::map<WsKey, WsBuf> g_ws;

// *from_pool: stream capture in progress -> the buffer is a cudaMallocAsync allocation of the capturing stream (a node of
// the graph); the caller frees it with cudaFreeAsync after the finalize kernel
static float *workspace(int dev, size_t bytes, cudaStream_t st, bool *from_pool) {
  *from_pool = false;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cs);
  if (cs != cudaStreamCaptureStatusNone) {
    void *p = nullptr;
    if (cudaMallocAsync(&p, bytes, st) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    *from_pool = true;
    return static_cast<float *>(p);
  }
  std::lock_guard<std::mutex> lk(g_ws_mu);
  WsBuf &b = g_ws[WsKey{dev, st}];
  if (b.bytes >= bytes) return b.p;
  if (b.p) cudaFree(b.p);
  const size_t want = bytes < (size_t)(64u << 20) ? (size_t)(64u << 20) : bytes;
  if (cudaMalloc(&b.p, want) != cudaSuccess) { b.p = nullptr; b.bytes = 0; cudaGetLastError(); return nullptr; }
  b.bytes = want;
  return b.p;
}
}  // namespace g4

static int ilog2_(int v) { int s = 0; while ((1 << s) < v) s++; return s; }

#include "gemm_4bit_small.cuh"
#include "gemm_4bit_wide.cuh"

// returns 0 ok, 1 shape not taken by the fused kernel (caller uses dequantize + matmul), 2 error
template <typename T>
int gemm_4bit(int batch, int N, int K, const T *A, const unsigned char *B, const float *absmax, const float *datatype,
              const T *bias, T *out, int blocksize) {
  using namespace g4;
  if (batch <= 0 || N <= 0) return 0;
  if (batch > kMaxNB || K < TK || (K % TK) != 0 || blocksize < 64 || (blocksize & (blocksize - 1)) != 0 ||
      (reinterpret_cast<uintptr_t>(A) % 16) != 0 || (reinterpret_cast<uintptr_t>(B) % 32) != 0 || ((K / 2) % 32) != 0)
    return 1;
  static int off = -1;
  if (off < 0) { const char *e = getenv("BNB_B200_GEMM4"); off = (e && e[0] == '0') ? 1 : 0; }
  if (off) return 1;
  int dev = 0;
  cudaGetDevice(&dev);
  dev = dev < 64 ? dev : 63;
  static int num_sms[64] = {0};
  if (!num_sms[dev]) cudaDeviceGetAttribute(&num_sms[dev], cudaDevAttrMultiProcessorCount, dev);
  cudaStream_t st = current_stream();
  static int small_off = -1;
  if (small_off < 0) { const char *e = getenv("BNB_B200_GEMM4_SMALL"); small_off = (e && e[0] == '0') ? 1 : 0; }
  if (batch <= 64 && !small_off && K / TK >= 8)
    return gemm_4bit_small<T>(batch, N, K, A, B, absmax, datatype, bias, out, ilog2_(blocksize), num_sms[dev], dev, st);

  static int wide_off = -1;
  if (wide_off < 0) { const char *e = getenv("BNB_B200_GEMM4_WIDE"); wide_off = (e && e[0] == '0') ? 1 : 0; }
  if (!wide_off && K / TK >= 4)
    return gemm_4bit_wide<T>(batch, N, K, A, B, absmax, datatype, bias, out, ilog2_(blocksize), num_sms[dev], dev, st);

  Args a{};
  a.batch = batch; a.N = N; a.K = K; a.blocksize = blocksize; a.bs_shift = ilog2_(blocksize);
  a.NB = (batch + 15) / 16 * 16;
  a.B = B; a.absmax = absmax; a.code = datatype; a.bias = bias; a.out = out;
  const int stage_bytes = kStageA + a.NB * 128;
  a.stages = kDqWarps / 4;                                 // one operand stage per dequant group (the groups' ring
                                                           // arithmetic needs stages >= groups)
  // ONE CTA per SM.  The packed-weight ring must hold a MULTIPLE of the group count of slots, so that a slot is always
  // drained by the same dequant group: TMA loads complete out of order, and a group waiting for round r + 1 of a slot
  // whose round r another group has not seen yet falls through the parity test (the phase two back has the same parity)
  // and dequantises the previous round's bytes.  Round 1's two-CTAs-per-SM configuration had 13 slots for 4 groups and
  // returned wrong results whenever the weights were not already in L2 (tools/gemm4_stress.py: 37 / 40 launches after
  // an L2 flush); its tests always ran on freshly written, L2-resident weights.  Batch <= 64 now takes k_gemm4_small,
  // this kernel keeps batch 65..256 and the shapes the small kernel refuses.
  const bool two_cta = false;
  a.wslots = (220 * 1024 - 2048 - a.stages * stage_bytes) / kStageW;
  if (a.wslots > kMaxWSlots) a.wslots = kMaxWSlots;
  a.wslots -= a.wslots % (kDqWarps / 4);
  if (a.wslots < a.stages) return 1;
  const int tiles = (N + TM - 1) / TM;
  const int target = num_sms[dev] * (two_cta ? 2 : 1);
  int splits = target / tiles;
  const int kblocks = K / TK;
  if (splits > kblocks / 8) splits = kblocks / 8;          // at least 8 stages of work per CTA
  if (splits > 16) splits = 16;
  if (splits < 1) splits = 1;
  const int kb_per = (kblocks + splits - 1) / splits;
  a.kper = kb_per * TK;
  a.splits = (kblocks + kb_per - 1) / kb_per;
  bool ws_from_pool = false;
  if (a.splits > 1) {
    a.ws = workspace(dev, (size_t)a.splits * batch * N * sizeof(float), st, &ws_from_pool);
    if (a.ws == nullptr) { a.splits = 1; a.kper = K; }
  }

  CUtensorMap tmX, tmW;
  if (!make_tmap_2d(&tmX, A, 2, (uint64_t)batch, (uint64_t)K, (uint32_t)a.NB, TK, true, std::is_same<T, __nv_bfloat16>::value) ||
      !make_tmap_2d(&tmW, B, 1, (uint64_t)N, (uint64_t)(K / 2), TM, TK / 2, false, false, false)) {
    if (ws_from_pool) cudaFreeAsync(a.ws, st);
    return 2;
  }
  const size_t smem = (size_t)a.stages * stage_bytes + (size_t)a.wslots * kStageW + 1024 /*align*/ + 1024 /*barriers, code*/;
  static bool attr_set[2][64] = {{false}};      // per element type AND per device
  const int ti = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
  if (!attr_set[ti][dev]) {
    ensure_max_dynamic_smem(reinterpret_cast<const void *>(k_gemm4_tcgen05<T>), 226 * 1024, "gemm_4bit smem attr");
    attr_set[ti][dev] = true;
  }
  k_gemm4_tcgen05<T><<<dim3(tiles, a.splits), kThreads, smem, st>>>(tmX, tmW, a);
  check_launch("gemm_4bit (tcgen05)");
  if (a.splits > 1) {
    const size_t total = (size_t)batch * N;
    int blocks = (int)((total + 255) / 256);
    if (blocks > num_sms[dev] * 8) blocks = num_sms[dev] * 8;
    k_gemm4_finalize<T><<<blocks, 256, 0, st>>>(a.ws, bias, out, a.splits, batch, N);
    check_launch("gemm_4bit (finalize)");
    if (ws_from_pool) cudaFreeAsync(a.ws, st);
  }
  return 0;
}

template int gemm_4bit<__half>(int, int, int, const __half *, const unsigned char *, const float *, const float *, const __half *, __half *, int);
template int gemm_4bit<__nv_bfloat16>(int, int, int, const __nv_bfloat16 *, const unsigned char *, const float *, const float *, const __nv_bfloat16 *, __nv_bfloat16 *, int);

}  // namespace bnb
