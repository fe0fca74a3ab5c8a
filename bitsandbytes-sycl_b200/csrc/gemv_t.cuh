// gemv_t.cuh -- k_gemv4_t: TMA-staged, warp-specialised batch-1 NF4/FP4 GEMV (included by gemv_4bit.cu).
//
// Replaces kgemm_4bit_inference_naive (reference kernel_gemm.cpp:1273-1388) and the two de-nesting launches in front
// of it (functional.py:1982-1984).  Arithmetic = the block-column scheme of gemv_4bit.cu (byte LUT -> mma.sync, column
// j of the accumulator = block j of a 512-element K chunk, absmax applied once per (row, block) in fp32).
//
// Roles inside the one CTA per SM:
//   * PRODUCER warp (the last one): one lane walks the CTA's (16-row tile, 512-K chunk) items in order and streams each
//     4 KB item with two TMA boxes (16 rows x 128 B, 128-byte swizzle, rows / columns past the matrix zero-filled by
//     the hardware) into a ring of up to 64 slots; full[slot] completes on the transaction bytes, empty[slot] is
//     armed by the consumer.  The ring -- not the register file -- holds the bytes in flight towards HBM (100+ KB per
//     SM), the producer never waits for the dependency of the launch (weights are constants) and the compute warps
//     spend no LSU wavefront and no register on global loads.
//   * W CONSUMER warps take items round-robin.  A lane lifts its weights out of the slot 16 bytes at a time, just
//     before the MMAs that use them (conflict-free LDS.128 through the swizzle), so a consumer needs ~8 weight
//     registers instead of 32-64: more warps fit, which is what hides the LUT-lookup latency.
//   * x: in REGISTERS when a warp always works on the same K chunk (K / 512 divides W; two-accumulator trick, see
//     gemv_v2.cuh), else predicated LDS.128 from shared memory.
//   * absmax: de-nested once per CTA in the prologue into fp32 rows in shared memory (ABS_SMEM; conflict-free
//     64-bit reads in the loop) when the CTA's rows fit, else per item from global memory as in round 1.
#pragma once

constexpr int kGtSlot = 4096;
constexpr int kGtMaxSlots = 64;
constexpr int kGtCode2 = 65536;                    // after the 64 KB byte LUT
constexpr int kGtBars = 65536 + 1024;              // full[64] | empty[64]
constexpr int kGtRing = 65536 + 2048;              // 1024-byte aligned slots

struct GemvTmaps { CUtensorMap m[4]; };

template <typename T, bool NESTED, int WARPS, bool XREG, bool ABS_SMEM, bool MULTI>
__global__ void __launch_bounds__((WARPS + 1) * 32, 1)
k_gemv4_t(const GemvArgs a, const __grid_constant__ GemvTmaps tm, int x_blocks_padded, int tiles_total, int abs_pitch, int nslots) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t raw_s = (uint32_t)__cvta_generic_to_shared(smem_raw);
  unsigned char *smem = smem_raw + (((raw_s + 1023u) & ~1023u) - raw_s);     // swizzled TMA destinations: 1024-byte aligned
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
  constexpr int CT = WARPS * 32;                   // consumer threads
  asm volatile("griddepcontrol.launch_dependents;");

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int kb = a.K >> 6;                 // blocks per row
  const int nch = (a.K + 511) >> 9;        // 512-element chunks per row (the last one may be half)
  const int t_begin = (int)(blockIdx.x * (unsigned)tiles_total / gridDim.x);
  const int t_end = (int)((blockIdx.x + 1) * (unsigned)tiles_total / gridDim.x);
  const int ntl = t_end - t_begin;
  const int ntl_max = (tiles_total + (int)gridDim.x - 1) / (int)gridDim.x;
  const int nitems = ntl * nch;

  float *s_code2 = reinterpret_cast<float *>(smem + kGtCode2);
  const uint32_t full_s = smem_base + kGtBars, empty_s = full_s + kGtMaxSlots * 8;
  const uint32_t ring_s = smem_base + kGtRing;
  float *s_abs = reinterpret_cast<float *>(smem + kGtRing + (size_t)nslots * kGtSlot);
  unsigned char *s_x = reinterpret_cast<unsigned char *>(s_abs + (ABS_SMEM ? (size_t)ntl_max * 16 * abs_pitch : 0));
  float *s_part = reinterpret_cast<float *>(s_x + (XREG ? 0 : (size_t)x_blocks_padded * kBcXPitch));   // [tile_local][warp][16]
  const uint32_t x_s = smem_base + (uint32_t)(s_x - smem);

  auto mat_of = [&](int tile) { return MULTI ? (int)(tile >= a.mt[1]) + (int)(tile >= a.mt[2]) + (int)(tile >= a.mt[3]) : 0; };
#define BNB_MSEL(arr, m) ((m) == 0 ? a.arr[0] : (m) == 1 ? a.arr[1] : (m) == 2 ? a.arr[2] : a.arr[3])

  if (tid == 0) {
    for (int i = 0; i < nslots; i++) { tc::mbar_init(full_s + i * 8, 1); tc::mbar_init(empty_s + i * 8, 1); }
    tc::fence_barrier_init();
  }
  __syncthreads();

  // ------------------------------------------------------------------------------------------ producer
  if (warp == WARPS) {
    // every lane streams its own items (lane, lane + 32, ...): the waits and the TMA issues of 32 items overlap; one
    // lane walking all items serially needs ~0.2 us per item and starves the consumers (measured)
    if (!MULTI && lane == 0) tc::prefetch_tmap(&tm.m[0]);
    const int nprod = nslots < 32 ? nslots : 32;
    if (lane < nprod) {
      for (int i = lane; i < nitems; i += nprod) {
        const int slot = i % nslots, round = i / nslots;
        const int tl_ = i / nch, c_ = i - tl_ * nch;
        const uint32_t bar = full_s + slot * 8, dst = ring_s + slot * kGtSlot;
        if (round > 0) tc::mbar_wait(empty_s + slot * 8, (uint32_t)(round & 1) ^ 1u);   // released by the consumer of the previous round
        const int tile = t_begin + tl_;
        const int m = mat_of(tile);
        const int lt = MULTI ? tile - BNB_MSEL(mt, m) : tile;
        const CUtensorMap *map = &tm.m[m];
        tc::mbar_arrive_expect_tx(bar, kGtSlot);
        tc::tma_load_2d(dst, map, bar, c_ * 256, lt * 16);
        tc::tma_load_2d(dst + 2048, map, bar, c_ * 256 + 128, lt * 16);
      }
    }
    return;
  }

  // ------------------------------------------------------------------------------------------ consumers
  auto csync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(CT) : "memory"); };
  // table constants: one round of small loads
  const float c2v = (NESTED && tid < 256) ? __ldg(a.code2 + tid) : 0.f;
  float cv[16];
  float clo;
  if (a.tables_in_args == 2) {     // the host verified code == the NF4 table: immediates
    constexpr float nf4[16] = BNB_NF4_TABLE;
#pragma unroll
    for (int u = 0; u < 16; u++) cv[u] = nf4[u];
    clo = nf4[0];
#pragma unroll
    for (int u = 1; u < 16; u++) clo = (((tid >> 3) & 15) == u) ? nf4[u] : clo;
  } else {
    if ((reinterpret_cast<uintptr_t>(a.code) & 15) == 0) {
      const float4 *cg = reinterpret_cast<const float4 *>(a.code);
#pragma unroll
      for (int u = 0; u < 4; u++) { const float4 f = __ldg(cg + u); cv[4 * u] = f.x; cv[4 * u + 1] = f.y; cv[4 * u + 2] = f.z; cv[4 * u + 3] = f.w; }
    } else {
#pragma unroll
      for (int u = 0; u < 16; u++) cv[u] = __ldg(a.code + u);
    }
    clo = __ldg(a.code + ((tid >> 3) & 15));
  }

  // absmax of the CTA's rows (ABS_SMEM): groups of 4 consecutive blocks of a row (kb % 4 == 0)
  const int total4 = ABS_SMEM ? ntl * 4 * kb : 0;
  struct AbsG { uint32_t q; float am2; float off; float4 f; int dst; };
  auto abs_load = [&](AbsG &d, int gi) {
    d.dst = -1; d.q = 0; d.am2 = 0.f; d.off = a.offset; d.f = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gi >= total4) return;
    const int v = gi * 4;
    const int rowlin = v / kb, blk = v - rowlin * kb;
    const int tile = t_begin + (rowlin >> 4);
    const int m = mat_of(tile);
    const int lt = MULTI ? tile - BNB_MSEL(mt, m) : tile;
    const int Nm = MULTI ? BNB_MSEL(mN, m) : a.N;
    const int grow = lt * 16 + (rowlin & 15);
    d.dst = rowlin * abs_pitch + blk;
    if (grow >= Nm) return;                       // rows past N: zeros (their outputs are never stored)
    const size_t idx = (size_t)grow * kb + blk;
    if (NESTED) {
      const unsigned char *qm = MULTI ? BNB_MSEL(mq, m) : a.qabsmax;
      const float *am2m = MULTI ? BNB_MSEL(mam2, m) : a.absmax2;
      d.q = __ldg(reinterpret_cast<const uint32_t *>(qm + idx));
      d.am2 = __ldg(am2m + (idx >> a.bs2_shift));
      if (MULTI) d.off = BNB_MSEL(moff, m);
    } else {
      d.f = __ldg(reinterpret_cast<const float4 *>(a.absmax + idx));
    }
  };
  auto abs_store = [&](const AbsG &d) {
    if (d.dst < 0) return;
    float4 f = d.f;
    if (NESTED) {
      f.x = __fadd_rn(__fmul_rn(s_code2[d.q & 0xFFu], d.am2), d.off);
      f.y = __fadd_rn(__fmul_rn(s_code2[(d.q >> 8) & 0xFFu], d.am2), d.off);
      f.z = __fadd_rn(__fmul_rn(s_code2[(d.q >> 16) & 0xFFu], d.am2), d.off);
      f.w = __fadd_rn(__fmul_rn(s_code2[d.q >> 24], d.am2), d.off);
    }
    *reinterpret_cast<float4 *>(s_abs + d.dst) = f;
  };
  AbsG ag[2];
  abs_load(ag[0], tid);
  abs_load(ag[1], tid + CT);

  // byte LUT: e -> {T(code[e >> 4]), T(code[e & 15])}, replicated for the 32 lanes (bank == lane): 8 threads write one
  // 128-byte entry with conflict-free 128-bit stores
  {
    const int j8 = tid & 7;
    const uint32_t lo16 = MmaT<T>::pack(clo, 0.0f) << 16;
    constexpr int EPI = CT / 8;                      // entries per iteration
    static_assert(EPI % 16 == 0, "consumer warps must be a multiple of 4");
    const int hsel = (tid >> 3) >> 4;
#pragma unroll
    for (int it = 0; it < (256 + EPI - 1) / EPI; it++) {
      const int e = (tid >> 3) + it * EPI;
      constexpr int HB = EPI / 16;
      float chi = cv[it * HB < 15 ? it * HB : 15];
#pragma unroll
      for (int h = 1; h < HB; h++) chi = (hsel == h) ? cv[it * HB + h < 15 ? it * HB + h : 15] : chi;
      const uint32_t v = (MmaT<T>::pack(chi, 0.0f) & 0xFFFFu) | lo16;
      if (e < 256) *reinterpret_cast<uint4 *>(smem + e * 256 + j8 * 16) = make_uint4(v, v, v, v);
    }
  }
  if (NESTED && tid < 256) s_code2[tid] = c2v;
  for (int i = tid; i < ntl * WARPS * 16; i += CT) s_part[i] = 0.f;
  csync();                                                          // code2 table visible
  if (ABS_SMEM) {
    abs_store(ag[0]);
    abs_store(ag[1]);
    for (int base = tid + 2 * CT; base < total4; base += 2 * CT) {
      abs_load(ag[0], base);
      abs_load(ag[1], base + CT);
      abs_store(ag[0]);
      abs_store(ag[1]);
    }
  }

  asm volatile("griddepcontrol.wait;" ::: "memory");               // x (and out) belong to the previous kernel until here
  if (a.sig_local != nullptr && a.do_wait) {
    if (tid < a.npeers) {
      const unsigned int target = ld_acquire_sys(a.epoch) * (unsigned int)a.ngroups + (unsigned int)a.gidx;
      const unsigned int *slot = a.sig_local + tid;
      const long long t0 = clock64();
      while ((int)(ld_acquire_sys(slot) - target) < 0) {
        if (clock64() - t0 > 4000000000ll) __trap();
      }
    }
    csync();
  }

  int tl = 0, c = warp;                // compute cursor: item `warp`, then + WARPS
  while (c >= nch) { c -= nch; tl++; }
  auto advance = [&](int &tl_, int &c_) {
    c_ += WARPS;
    while (c_ >= nch) { c_ -= nch; tl_++; }
  };

  // ---- x
  uint32_t xr[XREG ? 32 : 1];
  if (XREG) {
    // this warp's chunk never changes (nch divides WARPS): lanes (g == t) and (g == t + 4) keep block g of the chunk
    const bool feeds = (g == t) || (g == t + 4);
    const int xb = c * 8 + g;
    const uint4 *xg = reinterpret_cast<const uint4 *>(a.x) + (size_t)xb * 8;
#pragma unroll
    for (int u = 0; u < 8; u++) {
      uint4 v = make_uint4(0, 0, 0, 0);
      if (feeds && xb < kb && tl < ntl) v = ld_x_u4(xg + u);
      xr[(4 * u) % (XREG ? 32 : 1)] = v.x; xr[(4 * u + 1) % (XREG ? 32 : 1)] = v.y;
      xr[(4 * u + 2) % (XREG ? 32 : 1)] = v.z; xr[(4 * u + 3) % (XREG ? 32 : 1)] = v.w;
    }
  } else {
    const uint4 *xg = reinterpret_cast<const uint4 *>(a.x);
    const int pieces = x_blocks_padded * 8, valid = a.K >> 3;
    for (int p0 = tid; p0 < pieces; p0 += 4 * CT) {   // four independent loads in flight per thread
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int p = p0 + u * CT;
        v[u] = p < valid ? ld_x_u4(xg + p) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int p = p0 + u * CT;
        if (p < pieces) *reinterpret_cast<uint4 *>(s_x + (p >> 3) * kBcXPitch + (p & 7) * 16) = v[u];
      }
    }
  }
  csync();                                                          // absmax rows (and x) visible

  const uint32_t lane4 = (uint32_t)(lane * 4);
  const uint32_t act0 = (g == t), act1 = (g == t + 4);
  const uint32_t xlane = x_s + g * kBcXPitch;
  // this lane's two 16-byte pieces (q = 0, 1) of block t inside a swizzled 128-byte row: chunk (2t + q) ^ (row & 7)
  const uint32_t wl0 = (uint32_t)(g * 128 + (((2 * t) ^ g) << 4));
  const uint32_t wl1 = (uint32_t)(g * 128 + (((2 * t + 1) ^ g) << 4));
  uint32_t b0[4] = {0, 0, 0, 0}, b1[4] = {0, 0, 0, 0};
  float acc0 = 0.f, acc1 = 0.f;
  int slot = warp % nslots;
  uint32_t parity = (uint32_t)(warp / nslots) & 1u;

  while (tl < ntl) {
    int ntl_ = tl, nc = c;
    advance(ntl_, nc);
    // the four absmax values of this lane's partial sums: rows g, g + 8; blocks 2t, 2t + 1 of the chunk
    float2 am_lo, am_hi;
    if (ABS_SMEM) {
      const float *ap = s_abs + (tl * 16 + g) * abs_pitch + min(c * 8 + 2 * t, kb - 2);
      am_lo = *reinterpret_cast<const float2 *>(ap);
      am_hi = *reinterpret_cast<const float2 *>(ap + 8 * abs_pitch);
    } else {
      const int tile = t_begin + tl;
      const int m = mat_of(tile);
      const int lt = MULTI ? tile - BNB_MSEL(mt, m) : tile;
      const int Nm = MULTI ? BNB_MSEL(mN, m) : a.N;
      const float off = MULTI ? BNB_MSEL(moff, m) : a.offset;
      float2 am[2];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int row = min(lt * 16 + g + 8 * h, Nm - 1);
        const size_t idx = (size_t)row * kb + min(c * 8 + 2 * t, kb - 2);
        if (NESTED) {
          const unsigned char *qm = MULTI ? BNB_MSEL(mq, m) : a.qabsmax;
          const float *am2m = MULTI ? BNB_MSEL(mam2, m) : a.absmax2;
          const uint32_t q2 = __ldg(reinterpret_cast<const unsigned short *>(qm + idx));
          const float am2 = __ldg(am2m + (idx >> a.bs2_shift));
          am[h].x = __fadd_rn(__fmul_rn(s_code2[q2 & 0xFFu], am2), off);
          am[h].y = __fadd_rn(__fmul_rn(s_code2[q2 >> 8], am2), off);
        } else {
          am[h] = __ldg(reinterpret_cast<const float2 *>(a.absmax + idx));
        }
      }
      am_lo = am[0]; am_hi = am[1];
    }
    const uint32_t slot_s = ring_s + slot * kGtSlot;
    const uint32_t xc = xlane + c * (8 * kBcXPitch);
    float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
    tc::mbar_wait(full_s + slot * 8, parity);
#pragma unroll
    for (int j = 0; j < 2; j++) {          // j = 0: blocks 0-3 (columns 0-3), j = 1: blocks 4-7 (columns 4-7)
#pragma unroll
      for (int q = 0; q < 2; q++) {        // 16 bytes of each of the lane's two rows = 4 x 2 MMAs
        const uint4 v0 = lds_u128(slot_s + j * 2048 + (q ? wl1 : wl0));
        const uint4 v1 = lds_u128(slot_s + j * 2048 + 1024 + (q ? wl1 : wl0));
        const uint32_t s0a[4] = {v0.x, v0.y, v0.z, v0.w}, s1a[4] = {v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int wi = 0; wi < 4; wi++) {
          const int mg = 4 * q + wi;
          if (!XREG) {
            if (j == 0) lds_x4_pred(b0, xc + mg * 16, act0);
            else lds_x4_pred(b1, xc + mg * 16, act1);
          }
          const uint32_t s0 = s0a[wi], s1 = s1a[wi];
#pragma unroll
          for (int mm = 0; mm < 2; mm++) {
            const uint32_t selA = 0x7604u | ((2 * mm) << 4), selB = 0x7604u | ((2 * mm + 1) << 4);
            uint32_t af[4];
            af[0] = *reinterpret_cast<const uint32_t *>(smem + __byte_perm(s0, lane4, selA));
            af[1] = *reinterpret_cast<const uint32_t *>(smem + __byte_perm(s1, lane4, selA));
            af[2] = *reinterpret_cast<const uint32_t *>(smem + __byte_perm(s0, lane4, selB));
            af[3] = *reinterpret_cast<const uint32_t *>(smem + __byte_perm(s1, lane4, selB));
            if (XREG) {
              if (j == 0) MmaT<T>::mma(d0, af, xr[(4 * mg + 2 * mm) % (XREG ? 32 : 1)], xr[(4 * mg + 2 * mm + 1) % (XREG ? 32 : 1)]);
              else MmaT<T>::mma(d1, af, xr[(4 * mg + 2 * mm) % (XREG ? 32 : 1)], xr[(4 * mg + 2 * mm + 1) % (XREG ? 32 : 1)]);
            } else {
              if (j == 0) MmaT<T>::mma(d0, af, b0[2 * mm], b0[2 * mm + 1]);
              else MmaT<T>::mma(d0, af, b1[2 * mm], b1[2 * mm + 1]);
            }
          }
        }
      }
    }
    // every lane has lifted its bytes (their LDS results were consumed above): hand the slot back to the producer
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(empty_s + slot * 8);
    slot += WARPS;
    while (slot >= nslots) { slot -= nslots; parity ^= 1u; }

    if (XREG) {   // columns 0-3 of d0 and 4-7 of d1 are the block sums (gemv_v2.cuh)
      const bool lo = t < 2;
      d0[0] = lo ? d0[0] : d1[0]; d0[1] = lo ? d0[1] : d1[1]; d0[2] = lo ? d0[2] : d1[2]; d0[3] = lo ? d0[3] : d1[3];
    }
    acc0 = __fmaf_rn(d0[0], am_lo.x, acc0);
    acc0 = __fmaf_rn(d0[1], am_lo.y, acc0);
    acc1 = __fmaf_rn(d0[2], am_hi.x, acc1);
    acc1 = __fmaf_rn(d0[3], am_hi.y, acc1);
    if (ntl_ != tl) {   // this warp is done with the tile: park its partial sums
      acc0 += __shfl_xor_sync(0xffffffffu, acc0, 1);
      acc1 += __shfl_xor_sync(0xffffffffu, acc1, 1);
      acc0 += __shfl_xor_sync(0xffffffffu, acc0, 2);
      acc1 += __shfl_xor_sync(0xffffffffu, acc1, 2);
      if (t == 0) {
        float *slotp = s_part + (tl * WARPS + warp) * 16;
        slotp[g] = acc0;
        slotp[g + 8] = acc1;
      }
      acc0 = acc1 = 0.f;
    }
    tl = ntl_; c = nc;
  }
  csync();
  for (int i = tid; i < ntl * 16; i += CT) {
    const int tile_l = i >> 4, row = i & 15;
    const float *p = s_part + tile_l * WARPS * 16 + row;
    float sum = 0.f;
#pragma unroll
    for (int wq = 0; wq < WARPS; wq++) sum += p[wq * 16];
    if (MULTI) {
      const int m = mat_of(t_begin + tile_l);
      const int rr = (t_begin + tile_l - BNB_MSEL(mt, m)) * 16 + row;
      if (rr < BNB_MSEL(mN, m)) {
        const T v = from_float<T>(sum);
        reinterpret_cast<T *>(BNB_MSEL(mout, m))[rr] = v;
#pragma unroll
        for (int pr = 0; pr < 7; pr++)                                       // NVLink P2P stores
          if (pr < a.npeers)
            reinterpret_cast<T *>(m == 0 ? a.mpeer[0][pr] : m == 1 ? a.mpeer[1][pr] : m == 2 ? a.mpeer[2][pr] : a.mpeer[3][pr])[rr] = v;
      }
      continue;
    }
    const int r = (t_begin + tile_l) * 16 + row;
    if (r < a.N) {
      const T v = from_float<T>(sum);
      reinterpret_cast<T *>(a.out)[r] = v;
      for (int pr = 0; pr < a.npeers; pr++) reinterpret_cast<T *>(a.peer_out[pr])[r] = v;   // NVLink P2P stores
    }
  }
  if (a.sig_local != nullptr && a.do_signal) {
    __threadfence_system();                       // this thread's peer stores are visible system-wide
    csync();
    if (tid == 0) {
      const unsigned int done = atomicAdd(a.cta_counter, 1u);
      if (done == gridDim.x - 1) {                // last CTA of the grid: the whole slice is out
        *a.cta_counter = 0u;
        __threadfence_system();
        const unsigned int seq = ld_acquire_sys(a.epoch) * (unsigned int)a.ngroups + (unsigned int)a.gidx + 1u;
        for (int pr = 0; pr < a.npeers; pr++) st_release_sys(a.sig_peer[pr], seq);
      }
    }
  }
#undef BNB_MSEL
}

// host side: pick the variant; false = shape not taken (caller falls back to the register-ring kernel)
template <typename T, bool NESTED, bool MULTI>
static bool launch_t(const GemvArgs &a, int tiles, int sms) {
  static int warps_env = -1, pdl_off = 0, xreg_env = 1, slots_env = 0;
  if (warps_env < 0) {
    const char *e = getenv("BNB_B200_GEMV_TW"); warps_env = e ? atoi(e) : 0;
    const char *f = getenv("BNB_B200_GEMV_PDL"); pdl_off = (f && f[0] == '0') ? 1 : 0;
    const char *x = getenv("BNB_B200_GEMV_XREG"); xreg_env = (x && x[0] == '0') ? 0 : 1;
    const char *s = getenv("BNB_B200_GEMV_SLOTS"); slots_env = s ? atoi(s) : 0;
  }
  const int kb = a.K / 64, nch = ceil_div(a.K, 512);
  const int abs_pitch = ceil_div(kb, 32) * 32 + 8;
  const int xblocks = nch * 8;
  int warps = warps_env ? warps_env : 16;
#ifdef BNB_GEMV_SWEEP
  if (warps != 16 && warps != 20 && warps != 24 && warps != 28) warps = 16;
#else
  if (warps != 16 && warps != 24) warps = 16;
#endif
  const bool xreg = xreg_env && (a.K % 512 == 0) && (warps % nch == 0) && (warps == 16 || warps == 24);
  const int grid = tiles < sms ? tiles : sms;
  const int ntl_max = ceil_div(tiles, grid);
  const size_t abs_bytes = (size_t)ntl_max * 16 * abs_pitch * 4;
  const size_t x_bytes = xreg ? 0 : (size_t)xblocks * kBcXPitch;
  const size_t part_bytes = (size_t)ntl_max * warps * 16 * sizeof(float);
  const size_t budget = (size_t)kBcSmemMax - 1024 - kGtRing;      // 1 KB of alignment slack
  if (x_bytes + part_bytes + (size_t)warps * kGtSlot > budget) return false;
  // absmax rows in shared memory when they leave room for a ring of at least two items per warp
  const bool abs_smem = abs_bytes + x_bytes + part_bytes + (size_t)2 * warps * kGtSlot <= budget;
  const size_t fixed = (abs_smem ? abs_bytes : 0) + x_bytes + part_bytes;
  int nslots = (int)((budget - fixed) / kGtSlot);
  if (nslots > kGtMaxSlots) nslots = kGtMaxSlots;
  if (slots_env > 0 && nslots > slots_env) nslots = slots_env;
  if (nslots < warps) return false;
  GemvTmaps tm;
  const int nmat = MULTI ? a.nmat : 1;
  for (int i = 0; i < 4; i++) {
    const int im = i < nmat ? i : 0;
    const unsigned char *Bm = MULTI ? a.mB[im] : a.B;
    const int Nm = MULTI ? a.mN[im] : a.N;
    if (i < nmat || i == 0) {
      if ((reinterpret_cast<uintptr_t>(Bm) % 16) != 0) return false;
      if (!make_tmap_2d(&tm.m[i], Bm, 1, (uint64_t)Nm, (uint64_t)(a.K / 2), 16, 128, false, false, true)) return false;
    } else {
      tm.m[i] = tm.m[0];
    }
  }
  const size_t need = (size_t)kGtRing + (size_t)nslots * kGtSlot + fixed + 1024;
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(grid); lc.blockDim = dim3((warps + 1) * 32); lc.dynamicSmemBytes = need; lc.stream = current_stream();
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr; lc.numAttrs = pdl_off ? 0 : 1;
#define GT_LAUNCH(WARPS_, XREG_, ABS_)                                                                                   \
  do {                                                                                                                  \
    auto kfn = k_gemv4_t<T, NESTED, WARPS_, XREG_, ABS_, MULTI>;                                                        \
    static bool attr_done = false;                                                                                      \
    if (!attr_done) { latch_error(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kBcSmemMax), "gemv t smem attr"); attr_done = true; } \
    latch_error(cudaLaunchKernelEx(&lc, kfn, a, tm, xblocks, tiles, abs_pitch, nslots), "gemv_4bit (tma) launch");      \
  } while (0)
#define GT_LAUNCH_W(WARPS_)                                                                                              \
  do {                                                                                                                  \
    if (xreg && abs_smem) GT_LAUNCH(WARPS_, true, true);                                                                \
    else if (xreg) GT_LAUNCH(WARPS_, true, false);                                                                      \
    else if (abs_smem) GT_LAUNCH(WARPS_, false, true);                                                                  \
    else GT_LAUNCH(WARPS_, false, false);                                                                               \
  } while (0)
#ifdef BNB_GEMV_SWEEP
  if (warps == 20) { if (abs_smem) GT_LAUNCH(20, false, true); else GT_LAUNCH(20, false, false); }
  else if (warps == 28) { if (abs_smem) GT_LAUNCH(28, false, true); else GT_LAUNCH(28, false, false); }
  else
#endif
  if (warps == 24) GT_LAUNCH_W(24);
  else GT_LAUNCH_W(16);
#undef GT_LAUNCH_W
#undef GT_LAUNCH
  check_launch("gemv_4bit (tma)");
  return true;
}
