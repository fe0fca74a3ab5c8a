// gemm_4bit_wide.cuh -- batch 65..256 route of the fused 4-bit GEMM, and every batch the small-batch kernel refuses
// (included by gemm_4bit.cu).
//
// The operand is the reference's own: w = T(fp32 code[q] * fp32 absmax), one rounding (kernel_quant.cpp:1449-1450), so it
// is bit-identical to what dequantize_4bit would have written, and ONE accumulator runs over all of K (at these widths
// reading per-block accumulators out of TMEM, what gemm_4bit_small.cuh does, would cost 2 * NB cycles per stage).
// Everything else is the machinery of the small-batch kernel:
//   * dequant = one PRMT + one conflict-free LDS.64 per packed byte out of a lane-replicated table of fp32 code PAIRS
//     (entry stride 256 B, 8 bytes per lane, on a 64 KB boundary of the shared window so that the PRMT result is the
//     address), two FMUL by absmax and one cvt.rn.{bf16x2,f16x2}.f32 per pair: ~2.7 instructions per weight element where
//     round 1's kernel (16-entry table, shifts and masks per nibble, STS.128 of a swizzled row) spent 9;
//   * the operand goes from registers straight to TENSOR MEMORY (tcgen05.st.32x32b.x32, thread = TMEM lane = weight
//     row) and the MMA takes A from TMEM; shared memory carries only the packed bytes, the table and the activations;
//   * the packed-weight ring is handed back after the stage is written and has a multiple of the group count of slots
//     (see gemm_4bit_small.cuh for the race this avoids), issue loops are warp-uniform with an elected lane.
// Warps: 0 packed-weight TMA | 1 MMA issue, TMEM owner | 2 activation TMA | 3 - | 4..15 dequant (three groups of four),
// which also run the epilogue.  One CTA per (128-row tile, K split); split-K partials go through the fp32 workspace.
#pragma once

namespace g4w {
constexpr int TM = 128, TK = 64;
constexpr int kStageW = TM * 32;           // 4 KB of packed weights per stage
constexpr int kWSlots = 12;
static_assert(kWSlots % 3 == 0 && kWSlots % 4 == 0, "a packed-weight slot must belong to one dequant group");
constexpr int kFirstDq = 4;
__host__ __device__ constexpr int threads_for(int G) { return (kFirstDq + 4 * G) * 32; }
constexpr int kLut = 65536;
// shared window: [base .. +48 KB) packed ring | barriers | table at the first 64 KB boundary | activation ring behind it
constexpr uint32_t kXRingMax = 96 * 1024;
constexpr int kSmemBytes = 223 * 1024;     // base 0x400: table at 0x10000, activation ring 0x20000 .. 0x38000
// activation ring: S * (NB / CG) * 128 <= 96 KB, S >= groups; a CTA pair (CG = 2) holds half of the batch rows per CTA
__host__ __device__ constexpr int stages_for(int NB, int CG) { return (CG == 2 || NB <= 192) ? 4 : 3; }
constexpr uint32_t kACol = 256;            // TMEM: accumulator in columns [0, NB), operand stages of 32 columns from 256

struct Args {
  int batch, N, K, bs_shift;
  int NB;            // UMMA N: batch rounded up to 16 (<= 256)
  int splits, kper;  // K elements per split (multiple of 64)
  const unsigned char *B;
  const float *absmax;
  const float *code;
  const void *bias;  // T[N] or null
  void *out;         // T[batch, N]           (splits == 1)
  float *ws;         // fp32 [splits, batch, N] (splits > 1)
  g4::OutSpec o;     // row stride of out, peer copies (N-sharded stacks)
};

__device__ __forceinline__ float2 lds64f(uint32_t saddr) {
  float2 v;
  asm("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(saddr));
  return v;
}

// CG = 2: a CTA pair (cluster of two) computes 256 weight rows; each CTA dequantises ITS 128 rows into its own TMEM and
// loads ITS half of the activation rows, the leader issues tcgen05.mma.cta_group::2 (M = 256).  Halves the activation
// traffic through shared memory, the estimated bound of the one-CTA kernel at NB >= 128.
template <typename T, int G, int CG, bool PUSH>   // G dequant groups of four warps (<= stages); PUSH: see gemm_4bit_small.cuh
__global__ void __launch_bounds__(threads_for(G), 1) k_gemm4_wide(const __grid_constant__ CUtensorMap tmX,
                                                          const __grid_constant__ CUtensorMap tmW, const __grid_constant__ Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int kDqWarps = 4 * G;
  const int NB = a.NB;
  const int S = stages_for(NB, CG);
  const int stageB = (NB / CG) * 128;        // this CTA's activation rows x 64 T, SWIZZLE_128B
  const int rank = CG == 2 ? (int)cg2::cta_rank() : 0;
  const uint32_t smem_base = tc::smem_u32(smem_raw);
  const uint32_t wring_s = (smem_base + 1023u) & ~1023u;
  const uint32_t bars_s = wring_s + kWSlots * kStageW;
  const uint32_t lut_s = (bars_s + 1024u + 0xFFFFu) & ~0xFFFFu;
  const uint32_t xring_s = lut_s + kLut;
  if (xring_s + kXRingMax > smem_base + (uint32_t)kSmemBytes) __trap();   // shared window base moved: layout no longer fits
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (bars_s - smem_base));
  uint64_t *full = bars, *done = bars + 4, *fullW = bars + 8, *emptyW = bars + 8 + kWSlots, *tfull = bars + 8 + 2 * kWSlots;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 9 + 2 * kWSlots);
  uint8_t *lut = smem_raw + (lut_s - smem_base);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * TM;
  const int split = blockIdx.y;
  const int k_begin = split * a.kper;
  const int k_end = min(a.K, k_begin + a.kper);
  const int nk = (k_end - k_begin) / TK;

  if (warp == 0 && lane == 0) tc::prefetch_tmap(&tmW);
  if (warp == 2 && lane == 0) tc::prefetch_tmap(&tmX);
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; s++) {
        tc::mbar_init(tc::smem_u32(full + s), 4 * CG + 1);   // dequant warps of the group (of both CTAs) + activation expect_tx
        tc::mbar_init(tc::smem_u32(done + s), 1);       // tcgen05.commit: operands consumed
      }
      for (int s = 0; s < kWSlots; s++) {
        tc::mbar_init(tc::smem_u32(fullW + s), 1);
        tc::mbar_init(tc::smem_u32(emptyW + s), 4);
      }
      tc::mbar_init(tc::smem_u32(tfull), 1);
      tc::fence_barrier_init();
    }
    __syncwarp();
    if (CG == 1) asm volatile("bar.sync 1, 96;" ::: "memory");   // the two TMA warps need only the mbarriers: they start now
    if (CG == 2) cg2::tmem_alloc(tc::smem_u32(tmem_slot), 512); else tc::tmem_alloc(tc::smem_u32(tmem_slot), 512);
  } else if (CG == 1 && (warp == 0 || warp == 2)) {
    asm volatile("bar.sync 1, 96;" ::: "memory");
  }
  if (warp >= kFirstDq) {
    // table e -> {code[e >> 4], code[e & 15]} (fp32 pair; the even element sits in the high nibble), one pair per lane
    const int dt = threadIdx.x - kFirstDq * 32;
    for (int idx = dt; idx < 256 * 16; idx += kDqWarps * 32) {
      const int e = idx >> 4;
      const float c0 = __ldg(a.code + (e >> 4)), c1 = __ldg(a.code + (e & 15));
      *reinterpret_cast<float4 *>(lut + e * 256 + (idx & 15) * 16) = make_float4(c0, c1, c0, c1);
    }
  }
  // everybody else also needs the table and the TMEM allocation: barrier 2 joins all warps but the two producers, whose
  // first loads are in flight while the table is being written
  uint32_t tmem_base = 0;
  if (CG == 2) {           // the peer's mbarriers must exist before anything signals them: one cluster barrier for everybody
    tc::fence_before_sync();
    cg2::cluster_sync();
    tc::fence_after_sync();
    tmem_base = *tmem_slot;
  } else if (warp != 0 && warp != 2) {
    tc::fence_before_sync();
    asm volatile("bar.sync 2, %0;" ::"r"((int)blockDim.x - 64) : "memory");
    tc::fence_after_sync();
    tmem_base = *tmem_slot;
  }

  if (warp == 0) {
    // ================= packed weights: TMA into the deep ring =================
    int wslot = 0; uint32_t wphase = 0;
    for (int i = 0; i < nk; i++) {
      tc::mbar_wait(tc::smem_u32(emptyW + wslot), wphase ^ 1);
      if (tc::elect_one()) {
        const uint32_t fw = tc::smem_u32(fullW + wslot);
        tc::mbar_arrive_expect_tx(fw, kStageW);
        tc::tma_load_2d(wring_s + wslot * kStageW, &tmW, fw, (k_begin + i * TK) >> 1, n0);
      }
      __syncwarp();
      if (++wslot == kWSlots) { wslot = 0; wphase ^= 1; }
    }
  } else if (warp == 2) {
    // ================= activations: TMA into the operand ring =================
    int stage = 0; uint32_t phase = 0;
    for (int kb = 0; kb < nk; kb++) {
      tc::mbar_wait(tc::smem_u32(done + stage), phase ^ 1);
      if (tc::elect_one()) {
        const uint32_t fb = tc::smem_u32(full + stage);
        if (CG == 2) {
          if (rank == 0) tc::mbar_arrive_expect_tx(fb, (uint32_t)(2 * stageB));   // both halves are counted on the leader's barrier
          cg2::tma_load_2d(xring_s + stage * stageB, &tmX, fb, k_begin + kb * TK, rank * (NB / 2));
        } else {
          tc::mbar_arrive_expect_tx(fb, (uint32_t)stageB);
          tc::tma_load_2d(xring_s + stage * stageB, &tmX, fb, k_begin + kb * TK, 0);
        }
      }
      __syncwarp();
      if (++stage == S) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: four k16 steps per stage into the one accumulator =================
    const uint32_t idesc = tc::umma_idesc(tc::kCFormatF32, std::is_same<T, __nv_bfloat16>::value ? 1u : 0u, TM * CG, (uint32_t)NB);
    int stage = 0; uint32_t phase = 0;
    for (int kb = 0; kb < (rank == 0 ? nk : 0); kb++) {     // the leader of a pair issues for both
      tc::mbar_wait(tc::smem_u32(full + stage), phase);
      tc::fence_after_sync();
      const uint64_t bdesc = tc::umma_desc_sw128_kmajor(xring_s + stage * stageB);
      const uint32_t ta = tmem_base + kACol + (uint32_t)(stage * 32);
      if (tc::elect_one()) {
#pragma unroll
        for (int k = 0; k < TK / 16; k++) {
          if (CG == 2) cg2::umma_f16_ts(tmem_base, ta + 8 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          else g4s::umma_f16_ts(tmem_base, ta + 8 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
        }
        if (CG == 2) cg2::umma_commit_both(tc::smem_u32(done + stage)); else tc::umma_commit(tc::smem_u32(done + stage));
        if (kb == nk - 1) { if (CG == 2) cg2::umma_commit_both(tc::smem_u32(tfull)); else tc::umma_commit(tc::smem_u32(tfull)); }
      }
      __syncwarp();
      if (++stage == S) { stage = 0; phase ^= 1; }
    }
  } else if (warp >= kFirstDq) {
    // ================= dequant producers: packed bytes -> T(code * absmax) pairs -> TMEM =================
    const int dt = threadIdx.x - kFirstDq * 32;
    const int r = dt & 127;                               // weight row inside the tile
    const int grp = dt >> 7;
    const int row = min(n0 + r, a.N - 1);                 // clamped: rows past N are computed and dropped
    const size_t ebase = (size_t)row * a.K + k_begin;
    const uint32_t lutlane = lut_s | (uint32_t)(lane * 8);   // PRMT operand b: bytes 0, 2, 3 of every lookup address
    const uint32_t ta_warp = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kACol;
    int stage = grp % S, wslot = grp % kWSlots;
    uint32_t phase = (uint32_t)(grp / S) & 1u, wphase = (uint32_t)(grp / kWSlots) & 1u;
    float am_next = grp < nk ? __ldg(a.absmax + ((ebase + (size_t)grp * TK) >> a.bs_shift)) : 0.f;
    for (int kb = grp; kb < nk; kb += G) {
      const float am = am_next;
      if (kb + G < nk) am_next = __ldg(a.absmax + ((ebase + (size_t)(kb + G) * TK) >> a.bs_shift));
      tc::mbar_wait(tc::smem_u32(fullW + wslot), wphase);
      const uint32_t wp = wring_s + wslot * kStageW + r * 32;
      const uint4 w0 = g4::lds128(wp), w1 = g4::lds128(wp + 16);
      const uint32_t w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const uint32_t wbar = tc::smem_u32(emptyW + wslot);  // handed back after the stage is written (gemm_4bit_small.cuh)
      wslot += G;
      while (wslot >= kWSlots) { wslot -= kWSlots; wphase ^= 1u; }

      tc::mbar_wait(tc::smem_u32(done + stage), phase ^ 1);
      tc::fence_after_sync();
      uint32_t o[32];                                     // word m = elements 2m (low half), 2m + 1 of this row's 64
#pragma unroll
      for (int c = 0; c < 8; c++)                         // one packed word = 8 elements
#pragma unroll
        for (int i = 0; i < 4; i++) {                     // byte i of the word
          const float2 cp = lds64f(__byte_perm(w[c], lutlane, 0x7604u | (i << 4)));
          o[c * 4 + i] = g4::pack2<T>(__fmul_rn(cp.x, am), __fmul_rn(cp.y, am));
        }
      g4s::tmem_st_32x32b_x32(ta_warp + (uint32_t)(stage * 32), o);
      g4s::tmem_st_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        tc::mbar_arrive(wbar);
        if (CG == 2) cg2::mbar_arrive_remote(tc::smem_u32(full + stage), 0); else tc::mbar_arrive(tc::smem_u32(full + stage));
      }
      stage += G;
      while (stage >= S) { stage -= S; phase ^= 1u; }
    }

    // ---- epilogue: TMEM lane quarter warp % 4, the three groups take alternate 32-column chunks
    const int q = warp & 3;
    const int orow = n0 + q * 32 + lane;
    tc::mbar_wait(tc::smem_u32(tfull), 0);
    tc::fence_after_sync();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float bias = 0.f;
    if (a.bias != nullptr && a.splits == 1 && orow < a.N) bias = to_float<T>(reinterpret_cast<const T *>(a.bias)[orow]);
    for (int c0 = grp * 32; c0 < NB; c0 += 32 * G) {
      uint32_t v[32];
      tc::tmem_ld_32x32b_x32(taddr + c0, v);              // columns beyond NB are never stored
      tc::tmem_ld_wait();
      if (orow < a.N) {
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const int b = c0 + j;
          if (b < a.batch) {
            const float acc = __uint_as_float(v[j]);
            if (a.splits == 1) {
              const T v = from_float<T>(__fadd_rn(acc, bias));
              if (PUSH) {
                reinterpret_cast<T *>(a.out)[(size_t)b * a.o.ldo + orow] = v;
#pragma unroll 1
                for (int p = 0; p < a.o.npeers; p++)                       // NVLink peer stores: the all-gather of an N-sharded stack
                  reinterpret_cast<T *>(a.o.peer[p])[(size_t)b * a.o.ldo + orow] = v;   // (param space, indexed in place: __grid_constant__)
              } else {
                reinterpret_cast<T *>(a.out)[(size_t)b * a.N + orow] = v;
              }
            }
            else a.ws[((size_t)split * a.batch + b) * a.N + orow] = acc;
          }
        }
      }
    }
  }

  tc::fence_before_sync();
  if (CG == 2) cg2::cluster_sync(); else __syncthreads();
  if (warp == 1) { if (CG == 2) cg2::tmem_dealloc(tmem_base, 512); else tc::tmem_dealloc(tmem_base, 512); }
}
}  // namespace g4w

// host: returns 0 ok, 2 error
template <typename T>
static int gemm_4bit_wide(int batch, int N, int K, const T *A, const unsigned char *B, const float *absmax, const float *datatype,
                          const T *bias, T *out, int bs_shift, int sms, int dev, cudaStream_t st, const g4::OutSpec &ospec) {
  using namespace g4w;
  Args a{};
  a.batch = batch; a.N = N; a.K = K; a.bs_shift = bs_shift;
  static int pair_min = -1;   // BNB_B200_GEMM4_WIDE_PAIR=<batch>: CTA pairs from this batch on (0: never)
  if (pair_min < 0) { const char *e = getenv("BNB_B200_GEMM4_WIDE_PAIR"); pair_min = e ? atoi(e) : 160; }
  const int tiles = (N + TM - 1) / TM;
  const bool pair = pair_min > 0 && batch >= pair_min && tiles >= 2;
  a.NB = pair ? (batch + 31) / 32 * 32 : (batch + 15) / 16 * 16;
  a.B = B; a.absmax = absmax; a.code = datatype; a.bias = bias; a.out = out; a.o = ospec;
  int splits = sms / ((tiles + 1) / 2 * 2);                                // K is split until the grid fills the SMs
  const int kblocks = K / TK;
  if (splits > kblocks / 8) splits = kblocks / 8;          // at least 8 stages of work per CTA
  if (splits > 16) splits = 16;
  if (splits < 1) splits = 1;
  const int kb_per = (kblocks + splits - 1) / splits;
  a.kper = kb_per * TK;
  a.splits = (kblocks + kb_per - 1) / kb_per;
  bool ws_from_pool = false;
  if (a.splits > 1) {
    a.ws = g4::workspace(dev, (size_t)a.splits * batch * N * sizeof(float), st, &ws_from_pool);
    if (a.ws == nullptr) { a.splits = 1; a.kper = K; }
  }
  CUtensorMap tmX, tmW;
  if (!make_tmap_2d(&tmX, A, 2, (uint64_t)batch, (uint64_t)K, (uint32_t)(pair ? a.NB / 2 : a.NB), TK, true, std::is_same<T, __nv_bfloat16>::value) ||
      !make_tmap_2d(&tmW, B, 1, (uint64_t)N, (uint64_t)(K / 2), TM, TK / 2, false, false, false)) {
    if (ws_from_pool) cudaFreeAsync(a.ws, st);
    return 2;
  }
  static int g_env = -1;   // BNB_B200_GEMM4_WIDE_G=3: three dequant groups at every width (A/B measurements)
  if (g_env < 0) { const char *e = getenv("BNB_B200_GEMM4_WIDE_G"); g_env = (e && e[0] == '3') ? 3 : 4; }
  const bool push = ospec.npeers > 0 || ospec.ldo != N;
  auto launch = [&](auto kfn, int G, bool cluster) {
    ensure_max_dynamic_smem(reinterpret_cast<const void *>(kfn), kSmemBytes, "gemm_4bit wide smem attr");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster ? (tiles + 1) / 2 * 2 : tiles, a.splits);
    cfg.blockDim = dim3(threads_for(G));
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = cluster ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kfn, tmX, tmW, a);
  };
  if (pair) {
    if (push) launch(k_gemm4_wide<T, 4, 2, true>, 4, true); else launch(k_gemm4_wide<T, 4, 2, false>, 4, true);
  } else if (stages_for(a.NB, 1) >= 4 && g_env == 4) {
    if (push) launch(k_gemm4_wide<T, 4, 1, true>, 4, false); else launch(k_gemm4_wide<T, 4, 1, false>, 4, false);
  } else {
    if (push) launch(k_gemm4_wide<T, 3, 1, true>, 3, false); else launch(k_gemm4_wide<T, 3, 1, false>, 3, false);
  }
  check_launch("gemm_4bit (wide, tcgen05)");
  if (a.splits > 1) {
    const size_t total = (size_t)batch * N;
    int blocks = (int)((total + 255) / 256);
    if (blocks > sms * 8) blocks = sms * 8;
    g4::k_gemm4_finalize<T><<<blocks, 256, 0, st>>>(a.ws, bias, out, a.splits, batch, N, ospec);
    check_launch("gemm_4bit (finalize)");
    if (ws_from_pool) cudaFreeAsync(a.ws, st);
  }
  return 0;
}
