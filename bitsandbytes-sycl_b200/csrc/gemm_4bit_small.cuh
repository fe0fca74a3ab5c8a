// gemm_4bit_small.cuh -- batch <= 32 route of the fused 4-bit GEMM (included by gemm_4bit.cu).
//
// At these widths the tile is HBM-bound and what the kernel spends is instructions per weight.  The per-block scale
// therefore leaves the per-weight path:
//   * the UMMA A operand is the UNSCALED code value T(code[q]): one PRMT + one conflict-free LDS per packed byte out of a
//     lane-replicated byte-pair table (64 KB, on a 64 KB boundary of the shared window so that the PRMT result IS the
//     address), 32 registers per thread and stage written straight to TENSOR MEMORY with one tcgen05.st.32x32b.x32
//     (thread = TMEM lane = weight row); the MMA takes A from TMEM -- one instruction per weight element in all;
//   * a stage is exactly one quantisation block of every row (64 k-elements, blocksize 64; a larger blocksize spans
//     whole stages), accumulated by four tcgen05.mma into its OWN slot of a ring of TMEM accumulators;
//   * "scaler" warps (thread = TMEM lane = weight row) lift a finished slot out of TMEM, multiply it by the fp32
//     absmax of (row, block) and add it to fp32 totals kept in registers: batch FMAs per row per block instead of 64
//     multiplies -- and code * absmax is never rounded to T, so the result is closer to the exact product than the
//     reference's dequantize-then-matmul (tests compare with both).
// Warps: 0 packed-weight TMA | 1, 3 MMA issue for even / odd stages | 2 activation TMA | 4..15 dequant (three groups of
// four) | 16.. scalers.  One mbarrier per stage collects everything the MMA needs, one tcgen05.commit per stage frees the
// operands and publishes the accumulator.  Persistent CTAs (one per SM) walk (128-row tile, K split) units; split-K
// partials go to the fp32 workspace of gemm_4bit.cu and are summed by k_gemm4_finalize in split order.
// Measurements, the role traces and the variants that lost are in profiles/r2_gemm4_small_roles.md.
#pragma once

namespace g4s {
constexpr int TM = 128, TK = 64;
constexpr int kStageW = TM * 32;           // 4 KB of packed weights per quantisation block of the tile
// dequant groups of four warps: group g fills stages g, g + 3, ...  (a fourth group, tried with an 8-deep stage ring,
// bought nothing: profiles/r2_gemm4_small_roles.md)
constexpr int kDqGroups = 3;
// packed-weight ring.  A multiple of the group count, so that a slot is always drained by the SAME group: TMA loads
// complete out of order, and a group that waited for round r + 1 of a slot another group has not yet seen round r of
// would fall through the parity test (the phase two back has the same parity) and read the previous round's bytes.
constexpr int kWBytes = 12 * 4096;           // packed-weight ring: 48 KB = 12 one-block or 6 two-block slots
__host__ __device__ constexpr int wslots_for(int KB) { return kWBytes / (4096 * KB); }
static_assert(wslots_for(1) % kDqGroups == 0 && wslots_for(2) % kDqGroups == 0, "a packed-weight slot must belong to one dequant group");
constexpr int kDqWarps = 4 * kDqGroups;
constexpr int kFirstDq = 4, kFirstSc = kFirstDq + kDqWarps;   // warps: 0 W-TMA | 1 MMA even | 2 X-TMA | 3 MMA odd | dequant | scalers
// scaler warps: one per TMEM lane quarter at NB = 16, two (each takes half of the accumulator columns) above
__host__ __device__ constexpr int scaler_warps(int NB) { return NB <= 16 ? 4 : 8; }
__host__ __device__ constexpr int threads_for(int NB) { return (kFirstSc + scaler_warps(NB)) * 32; }
constexpr int kLut = 65536;                // byte-pair table, entry stride 256 B, one word per lane
constexpr int kBarBytes = 512;
constexpr int kSmemBytes = 192 * 1024;     // rings below the second 64 KB boundary of the shared window, table above it
// stages = TMEM ring: S accumulator slots of NB columns + S weight stages of 32 columns <= 512 columns
// (KB = quantisation blocks per barrier round: every role's wait -> work -> arrive round trip covers KB blocks)
__host__ __device__ constexpr int stages_for(int NB, int KB) { return KB == 2 ? (NB <= 16 ? 5 : 4) : (NB <= 48 ? 6 : 5); }

struct Args {
  int batch, N, K, bs_shift;
  int NB;            // UMMA N: batch rounded up to 16 (16 or 32 as dispatched; the kernel itself is generic up to 64)
  int splits, kper;  // K elements per split (multiple of 64 * KB)
  int tiles;         // 128-row tiles
  const unsigned char *B;
  const float *absmax;
  const float *code;
  const void *bias;  // T[N] or null
  void *out;         // T[batch, N]           (splits == 1)
  float *ws;         // fp32 [splits, batch, N] (splits > 1)
  g4::OutSpec o;     // row stride of out, peer copies (N-sharded stacks)
#ifdef G4S_TRACE
  int dbg;           // trace build only: 1 skip the dequant work, 2 skip the MMAs, 4 skip the TMEM loads
#endif
};

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
        "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
        "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]: A is read from tensor memory -- row i in lane i, two 16-bit elements per column, eight
// columns per k16 step -- so the operand the dequant warps produce never goes through shared memory
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// explicit shared-space load: the table pointer is derived from the aligned dynamic-smem base, which the compiler would
// otherwise treat as a generic address (LD instead of LDS)
__device__ __forceinline__ uint32_t lds32(uint32_t saddr) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}

#ifdef G4S_TRACE
#define G4S_T0() const long long t_role0 = clock64(); long long t_w0 = 0, t_w1 = 0, t_w2 = 0
#define G4S_WAIT(acc, bar, par) do { const long long t_ = clock64(); tc::mbar_wait(bar, par); acc += clock64() - t_; } while (0)
#define G4S_DBG(bit) (a.dbg & (bit))
#define G4S_REPORT(name, cond) do { if (blockIdx.x == 5 && (cond)) printf("g4s %-8s warp %2d: total %lld wait0 %lld wait1 %lld wait2 %lld\n", name, (int)(threadIdx.x >> 5), clock64() - t_role0, t_w0, t_w1, t_w2); } while (0)
#else
#define G4S_T0() do {} while (0)
#define G4S_DBG(bit) false
#define G4S_WAIT(acc, bar, par) tc::mbar_wait(bar, par)
#define G4S_REPORT(name, cond) do {} while (0)
#endif

// One mbarrier per operand stage collects EVERYTHING the MMA issuer needs for that stage -- the four dequant warps of
// the group, the activation tile's TMA bytes, and the four scaler warps handing back the TMEM slot of the same index --
// so the single issuing thread pays one wait per stage (a try_wait costs ~100 cycles even when it succeeds, and three of
// them per 64-element stage were the critical path of the first version of this kernel).
// PUSH: the output has a row stride of its own and / or peer copies (N-sharded stacks).  A separate instantiation: with the
// peer pointers live across the main loops the plain kernel lost 5-25 % (14336x4096: batch 16 19.5 -> 20.8 us).
template <typename T, int NB16, int KB, bool PUSH>   // NB16 = NB / 16 (1..4); KB = 64-element blocks per round (2 only with NB <= 32)
__global__ void __launch_bounds__(threads_for(NB16 * 16), 1) k_gemm4_small(const __grid_constant__ CUtensorMap tmX,
                                                           const __grid_constant__ CUtensorMap tmW, const __grid_constant__ Args a) {
  constexpr int NB = NB16 * 16;
  constexpr int S = stages_for(NB, KB);
  constexpr int kWSlots = wslots_for(KB);
  constexpr int kSlotW = kStageW * KB;       // bytes of one packed-weight slot: 128 rows x 32 KB bytes
  static_assert(S * KB * (NB + 32) <= 512, "TMEM columns");
  constexpr int G = kDqGroups, SCW = scaler_warps(NB);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int tileB = NB * 128;            // activation tile of one block: NB rows x 64 T, SWIZZLE_128B
  constexpr int stageB = tileB * KB;
  constexpr uint32_t kAccCols = S * KB * NB; // TMEM: S x KB accumulator slots of NB columns, then S weight stages of 32 KB columns
  // activation ring | packed ring | barriers | ... | byte-pair table at the next 64 KB boundary OF THE SHARED WINDOW: with
  // bytes 0 and 1 of the table's address zero, one PRMT of (packed word, table | lane * 4) IS the lookup address -- no add
  const uint32_t smem_base = tc::smem_u32(smem_raw);
  const uint32_t ring_s0 = (smem_base + 1023u) & ~1023u;
  const uint32_t lut_s = (ring_s0 + (uint32_t)(S * stageB + kWBytes + kBarBytes) + 0xFFFFu) & ~0xFFFFu;
  if (lut_s + (uint32_t)kLut > smem_base + (uint32_t)kSmemBytes) __trap();   // shared window base moved: layout no longer fits
  uint8_t *ring = smem_raw + (ring_s0 - smem_base);
  uint8_t *wring = ring + S * stageB;
  uint8_t *lut = smem_raw + (lut_s - smem_base);
  uint64_t *bars = reinterpret_cast<uint64_t *>(wring + kWBytes);
  uint64_t *full = bars, *done = bars + 8, *fullW = bars + 16, *emptyW = bars + 16 + 12;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 16 + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int units = a.tiles * a.splits;

  if (warp == 0 && lane == 0) tc::prefetch_tmap(&tmW);
  if (warp == 2 && lane == 0) tc::prefetch_tmap(&tmX);
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; s++) {
        tc::mbar_init(tc::smem_u32(full + s), 4 + 1 + SCW);   // dequant warps + X expect_tx + scaler warps (TMEM slot free)
        tc::mbar_init(tc::smem_u32(done + s), 1);           // tcgen05.commit: operands consumed AND accumulator slot complete
      }
      for (int s = 0; s < kWSlots; s++) {
        tc::mbar_init(tc::smem_u32(fullW + s), 1);
        tc::mbar_init(tc::smem_u32(emptyW + s), 4);
      }
      tc::fence_barrier_init();
    }
    __syncwarp();
    asm volatile("bar.sync 1, 96;" ::: "memory");        // the two TMA warps need only the mbarriers: they start now
    tc::tmem_alloc(tc::smem_u32(tmem_slot), 512);
  } else if (warp == 0 || warp == 2) {
    asm volatile("bar.sync 1, 96;" ::: "memory");
  }
  if (warp >= kFirstDq && warp < kFirstSc) {
    // byte-pair table e -> {T(code[e >> 4]), T(code[e & 15])}, replicated for the 32 lanes: 8 threads per entry
    const int dt = threadIdx.x - kFirstDq * 32;
    const int j8 = dt & 7;
    for (int e = dt >> 3; e < 256; e += (kDqWarps * 32) / 8) {
      const uint32_t v = (g4::pack2<T>(__ldg(a.code + (e >> 4)), 0.f) & 0xFFFFu) | (g4::pack2<T>(__ldg(a.code + (e & 15)), 0.f) << 16);
      *reinterpret_cast<uint4 *>(lut + e * 256 + j8 * 16) = make_uint4(v, v, v, v);
    }
  }
  // everybody else also needs the table and the TMEM allocation: barrier 2 joins all warps but the two producers, whose
  // first loads are in flight while the table is being written
  uint32_t tmem_base = 0;
  if (warp != 0 && warp != 2) {
    tc::fence_before_sync();
    asm volatile("bar.sync 2, %0;" ::"r"((int)blockDim.x - 64) : "memory");
    tc::fence_after_sync();
    tmem_base = *tmem_slot;
  }

  auto unit_range = [&](int u, int &tile, int &split, int &k_begin, int &nk) {
    tile = u / a.splits;
    split = u - tile * a.splits;
    k_begin = split * a.kper;
    const int k_end = min(a.K, k_begin + a.kper);
    nk = (k_end - k_begin) / (TK * KB);   // rounds
  };

  if (warp == 0) {
    // ================= packed weights: TMA into the deep ring =================
    // The three single-issuer roles run their loops with the WHOLE warp (uniform control flow, uniform operands) and
    // elect one lane only around the issuing instructions: inside `if (lane == 0)` the compiler has to assume divergence
    // and wraps every UTMALDG / UTCHMMA / UTCBAR in an ELECT + R2UR.BROADCAST + BRA.U.ANY loop, ~60 cycles apiece --
    // with four MMAs and a commit per 64-element stage that loop WAS the critical path of this kernel.
    {
      int wslot = 0; uint32_t wphase = 0;
      G4S_T0();
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        int tile, split, k_begin, nk;
        unit_range(u, tile, split, k_begin, nk);
        for (int i = 0; i < nk; i++) {
          G4S_WAIT(t_w0, tc::smem_u32(emptyW + wslot), wphase ^ 1);
          if (tc::elect_one()) {
            const uint32_t fw = tc::smem_u32(fullW + wslot);
            tc::mbar_arrive_expect_tx(fw, kSlotW);
            tc::tma_load_2d(tc::smem_u32(wring + wslot * kSlotW), &tmW, fw, (k_begin + i * TK * KB) >> 1, tile * TM);
          }
          __syncwarp();
          if (++wslot == kWSlots) { wslot = 0; wphase ^= 1; }
        }
      }
      G4S_REPORT("W-TMA", lane == 0);
    }
  } else if (warp == 2) {
    // ================= activations: TMA into the operand ring =================
    {
      int stage = 0; uint32_t phase = 0;
      G4S_T0();
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        int tile, split, k_begin, nk;
        unit_range(u, tile, split, k_begin, nk);
        for (int kb = 0; kb < nk; kb++) {
          G4S_WAIT(t_w0, tc::smem_u32(done + stage), phase ^ 1);
          if (tc::elect_one()) {
            const uint32_t fb = tc::smem_u32(full + stage);
            tc::mbar_arrive_expect_tx(fb, (uint32_t)stageB);
#pragma unroll
            for (int blk = 0; blk < KB; blk++)
              tc::tma_load_2d(tc::smem_u32(ring + stage * stageB + blk * tileB), &tmX, fb, k_begin + (kb * KB + blk) * TK, 0);
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
      G4S_REPORT("X-TMA", lane == 0);
    }
  } else if (warp == 1 || warp == 3) {
    // ================= MMA issuers: four k16 steps per stage into the TMEM slot of the same index =================
    // Two warps, alternate stages: every stage has its own accumulator slot, so MMAs of different stages need no order
    // between them, and the ~350 cycles one thread spends per stage (wait, fence, issue, commit) run two abreast.
    {
      const int mine = warp >> 1;                          // warp 1: even items, warp 3: odd items
      int item = 0;
      const uint32_t idesc = tc::umma_idesc(tc::kCFormatF32, std::is_same<T, __nv_bfloat16>::value ? 1u : 0u, TM, (uint32_t)NB);
      int stage = 0; uint32_t phase = 0;
      G4S_T0();
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        int tile, split, k_begin, nk;
        unit_range(u, tile, split, k_begin, nk);
        for (int kb = 0; kb < nk; kb++, item++) {
          if ((item & 1) != mine) {
            if (++stage == S) { stage = 0; phase ^= 1; }
            continue;
          }
          G4S_WAIT(t_w0, tc::smem_u32(full + stage), phase);
          tc::fence_after_sync();
          const uint32_t ta = tmem_base + kAccCols + (uint32_t)(stage * 32 * KB);   // weights: 8 columns per k16 step
          if (tc::elect_one()) {                               // the same lane every time (commit tracks the issuing thread)
            if (!G4S_DBG(2))
#pragma unroll
            for (int blk = 0; blk < KB; blk++) {               // one accumulator slot per quantisation block
              const uint64_t bdesc = tc::umma_desc_sw128_kmajor(tc::smem_u32(ring + stage * stageB + blk * tileB));
#pragma unroll
              for (int k = 0; k < TK / 16; k++)
                umma_f16_ts(tmem_base + (uint32_t)((stage * KB + blk) * NB), ta + 32 * blk + 8 * k, bdesc + 2 * k, idesc, k != 0);
            }
            tc::umma_commit(tc::smem_u32(done + stage));
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
      G4S_REPORT("MMA", lane == 0);
    }
  } else if (warp >= kFirstDq && warp < kFirstSc) {
    // ================= dequant producers: packed bytes -> unscaled T(code) pairs =================
    const int dt = threadIdx.x - kFirstDq * 32;
    const int r = dt & 127;                               // weight row inside the tile
    const int grp = dt >> 7;
    const uint32_t lutlane = lut_s | (uint32_t)(lane * 4);   // PRMT operand b: bytes 0, 2, 3 of every lookup address
    const uint32_t wring_s = tc::smem_u32(wring);
    // TMEM lanes of this warp (a warp reaches lanes 32 * (warp % 4) ..., and row r = 32 * (warp % 4) + lane)
    const uint32_t ta_warp = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kAccCols;
    int stage = grp % S, wslot = grp % kWSlots;
    uint32_t phase = (uint32_t)(grp / S) & 1u, wphase = (uint32_t)(grp / kWSlots) & 1u;
    int total = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      int tile, split, k_begin, nk;
      unit_range(u, tile, split, k_begin, nk);
      total += nk;
    }
    G4S_T0();
    for (int it = grp; it < total; it += G) {
      G4S_WAIT(t_w0, tc::smem_u32(fullW + wslot), wphase);
      const uint32_t wp = wring_s + wslot * kSlotW + r * (32 * KB);
      uint32_t w[8 * KB];
#pragma unroll
      for (int h = 0; h < 2 * KB; h++) {
        const uint4 t = g4::lds128(wp + 16 * h);
        w[4 * h] = t.x; w[4 * h + 1] = t.y; w[4 * h + 2] = t.z; w[4 * h + 3] = t.w;
      }
      // The slot is handed back AFTER the stage has been written (below), not here: every table lookup consumes the
      // loaded registers, so by then both loads have returned.  An arrive issued right behind the LDS.128 (what this code
      // did first; an empty inline asm naming the registers does not make ptxas wait on the load scoreboard) let the
      // next round's TMA bytes land in rows a warp had not read yet -- seen only with cold caches, as a handful of rows
      // of one tile computed with the weights of the stage nine further on.
      const uint32_t wbar = tc::smem_u32(emptyW + wslot);
      wslot += G;
      while (wslot >= kWSlots) { wslot -= kWSlots; wphase ^= 1u; }

      G4S_WAIT(t_w1, tc::smem_u32(done + stage), phase ^ 1);
      tc::fence_after_sync();
#ifdef G4S_TRACE
      const long long t_d0 = clock64();
#endif
#pragma unroll
      for (int blk = 0; blk < KB; blk++) {
        uint32_t o[32];                                   // word m = elements 2m (low half), 2m + 1 of this row's 64
        if (!G4S_DBG(1))
#pragma unroll
        for (int c = 0; c < 8; c++)                       // one packed word = 8 elements
#pragma unroll
          for (int i = 0; i < 4; i++)                     // byte i of the word: even element in the high nibble = low half of the pair
            o[c * 4 + i] = lds32(__byte_perm(w[blk * 8 + c], lutlane, 0x7604u | (i << 4)));
        tmem_st_32x32b_x32(ta_warp + (uint32_t)(stage * 32 * KB + blk * 32), o);
      }
      tmem_st_wait();
#ifdef G4S_TRACE
      t_w2 += clock64() - t_d0;
#endif
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        tc::mbar_arrive(wbar);
        tc::mbar_arrive(tc::smem_u32(full + stage));
      }
      stage += G;
      while (stage >= S) { stage -= S; phase ^= 1u; }
    }
    G4S_REPORT("dequant", lane == 0);
  } else if (warp >= kFirstSc) {
    // ================= scalers: TMEM slot * absmax(row, block) -> fp32 totals; epilogue per unit =================
    const int q = warp & 3;                               // TMEM lane quarter of this warp
    const int part = (warp - kFirstSc) >> 2;              // which part of the NB accumulator columns
    constexpr int HC = NB / (SCW / 4);                    // columns (batch rows) this warp scales: 16, 24 or 32
    const int col0 = part * HC;
    const int r = q * 32 + lane;
    if (lane == 0)                                        // round 0: every TMEM slot starts free
      for (int s = 0; s < S; s++) tc::mbar_arrive(tc::smem_u32(full + s));
    int stage = 0; uint32_t phase = 0;
    G4S_T0();
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      int tile, split, k_begin, nk;
      unit_range(u, tile, split, k_begin, nk);
      const int orow = tile * TM + r;
      const int row = min(orow, a.N - 1);
      const size_t ebase = (size_t)row * a.K + k_begin;
      float tot[HC];
#pragma unroll
      for (int j = 0; j < HC; j++) tot[j] = 0.f;
      // absmax of (row, block), one round ahead
      float am_next[KB];
#pragma unroll
      for (int blk = 0; blk < KB; blk++) am_next[blk] = __ldg(a.absmax + ((ebase + (size_t)blk * TK) >> a.bs_shift));
      for (int kb = 0; kb < nk; kb++) {
        float am[KB];
#pragma unroll
        for (int blk = 0; blk < KB; blk++) am[blk] = am_next[blk];
        if (kb + 1 < nk && !G4S_DBG(8)) {
#pragma unroll
          for (int blk = 0; blk < KB; blk++) am_next[blk] = __ldg(a.absmax + ((ebase + (size_t)((kb + 1) * KB + blk) * TK) >> a.bs_shift));
        }
        G4S_WAIT(t_w0, tc::smem_u32(done + stage), phase);
        tc::fence_after_sync();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(stage * KB * NB + col0);
        // lift this warp's columns of the round's KB slots, hand them back, then scale
#ifdef G4S_TRACE
        const long long t_s0 = clock64();
#endif
        uint32_t v[KB][HC / 8][8];
        if (!G4S_DBG(4))
#pragma unroll
        for (int blk = 0; blk < KB; blk++)
#pragma unroll
          for (int c = 0; c < HC / 8; c++) tmem_ld_32x32b_x8(taddr + blk * NB + c * 8, v[blk][c]);
        tc::tmem_ld_wait();
#ifdef G4S_TRACE
        const long long t_s1 = clock64();
        t_w1 += t_s1 - t_s0;
#endif
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(tc::smem_u32(full + stage));   // the slots may be overwritten by this stage's next round
#pragma unroll
        for (int blk = 0; blk < KB; blk++)
#pragma unroll
          for (int c = 0; c < HC / 8; c++)
#pragma unroll
            for (int j = 0; j < 8; j++) tot[c * 8 + j] = __fmaf_rn(__uint_as_float(v[blk][c][j]), am[blk], tot[c * 8 + j]);
#ifdef G4S_TRACE
        asm volatile("" ::"f"(tot[0]), "f"(tot[HC - 1]));
        t_w2 += clock64() - t_s1;
#endif
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
      if (orow < a.N) {
        float bias = 0.f;
        if (a.bias != nullptr && a.splits == 1) bias = to_float<T>(reinterpret_cast<const T *>(a.bias)[orow]);
#pragma unroll
        for (int j = 0; j < HC; j++) {
          const int b = col0 + j;
          if (b < a.batch) {
            if (a.splits == 1) {
              const T v = from_float<T>(__fadd_rn(tot[j], bias));
              if (PUSH) {
                reinterpret_cast<T *>(a.out)[(size_t)b * a.o.ldo + orow] = v;
#pragma unroll 1
                for (int p = 0; p < a.o.npeers; p++)                       // NVLink peer stores: the all-gather of an N-sharded stack
                  reinterpret_cast<T *>(a.o.peer[p])[(size_t)b * a.o.ldo + orow] = v;   // (param space, indexed in place: __grid_constant__)
              } else {
                reinterpret_cast<T *>(a.out)[(size_t)b * a.N + orow] = v;
              }
            }
            else a.ws[((size_t)split * a.batch + b) * a.N + orow] = tot[j];
          }
        }
      }
    }
    G4S_REPORT("scaler", lane == 0);
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace g4s

// host: batch <= 32.  returns 0 ok, 2 error
template <typename T>
static int gemm_4bit_small(int batch, int N, int K, const T *A, const unsigned char *B, const float *absmax, const float *datatype,
                           const T *bias, T *out, int bs_shift, int sms, int dev, cudaStream_t st, const g4::OutSpec &ospec) {
  using namespace g4s;
  Args a{};
  a.batch = batch; a.N = N; a.K = K; a.bs_shift = bs_shift;
  a.NB = (batch + 15) / 16 * 16;
  a.B = B; a.absmax = absmax; a.code = datatype; a.bias = bias; a.out = out; a.o = ospec;
  a.tiles = (N + TM - 1) / TM;
#ifdef G4S_TRACE
  { const char *e = getenv("G4S_DBG"); a.dbg = e ? atoi(e) : 0; }
#endif
  // K splits: the makespan of `units` equal units on `sms` persistent CTAs, each unit = its stages + ~6 stages of fill /
  // drain / epilogue; at least 8 stages per unit
  // one quantisation block per barrier round; two (NB <= 32, K allowing: TMEM holds S x KB x (NB + 32) columns) on request
  static int kb_env = -1;   // BNB_B200_GEMM4_SMALL_KB=2: two blocks per round (measured slower: profiles/r2_gemm4_small_roles.md)
  if (kb_env < 0) { const char *e = getenv("BNB_B200_GEMM4_SMALL_KB"); kb_env = (e && e[0] == '2') ? 2 : 1; }
  const int KB = (a.NB <= 32 && (K / TK) % 2 == 0 && kb_env == 2) ? 2 : 1;
  const int kblocks = K / (TK * KB);   // rounds
  int best = 1;
  long best_cost = -1;
  for (int s = 1; s <= 16 && kblocks / s >= 8 / KB; s++) {
    const int kb_per = (kblocks + s - 1) / s;
    const int ns = (kblocks + kb_per - 1) / kb_per;
    const long waves = ((long)a.tiles * ns + sms - 1) / sms;
    const long cost = waves * (kb_per + 6 / KB);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
  }
  const int kb_per = (kblocks + best - 1) / best;
  a.kper = kb_per * TK * KB;
  a.splits = (kblocks + kb_per - 1) / kb_per;
  bool ws_from_pool = false;
  if (a.splits > 1) {
    a.ws = g4::workspace(dev, (size_t)a.splits * batch * N * sizeof(float), st, &ws_from_pool);
    if (a.ws == nullptr) { a.splits = 1; a.kper = K; }
  }
  CUtensorMap tmX, tmW;
  if (!make_tmap_2d(&tmX, A, 2, (uint64_t)batch, (uint64_t)K, (uint32_t)a.NB, TK, true, std::is_same<T, __nv_bfloat16>::value) ||
      !make_tmap_2d(&tmW, B, 1, (uint64_t)N, (uint64_t)(K / 2), TM, (uint32_t)(TK / 2 * KB), false, false, false)) {
    if (ws_from_pool) cudaFreeAsync(a.ws, st);
    return 2;
  }
  const int units = a.tiles * a.splits;
  const int grid = units < sms ? units : sms;
  const size_t smem = kSmemBytes;
#define G4S_LAUNCH(NB16_, KB_, PUSH_)                                                                                     \
  do {                                                                                                                  \
    auto kfn = k_gemm4_small<T, NB16_, KB_, PUSH_>;                                                                     \
    ensure_max_dynamic_smem(reinterpret_cast<const void *>(kfn), kSmemBytes, "gemm_4bit small smem attr");             \
    kfn<<<grid, threads_for(NB16_ * 16), smem, st>>>(tmX, tmW, a);                                                     \
  } while (0)
  const bool push = ospec.npeers > 0 || ospec.ldo != N;
  if (push) {
    if (a.NB == 16) { if (KB == 2) G4S_LAUNCH(1, 2, true); else G4S_LAUNCH(1, 1, true); }
    else { if (KB == 2) G4S_LAUNCH(2, 2, true); else G4S_LAUNCH(2, 1, true); }
  } else {
    if (a.NB == 16) { if (KB == 2) G4S_LAUNCH(1, 2, false); else G4S_LAUNCH(1, 1, false); }
    else { if (KB == 2) G4S_LAUNCH(2, 2, false); else G4S_LAUNCH(2, 1, false); }
  }
#undef G4S_LAUNCH
  check_launch("gemm_4bit (small batch, tcgen05)");
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
