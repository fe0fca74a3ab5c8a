// gemv_v2.cuh -- k_gemv4_v2: the default batch-1 NF4/FP4 GEMV (included by gemv_4bit.cu).
//
// Same arithmetic as the block-column kernel of round 1 (byte LUT -> mma.sync, column j of the accumulator = block j
// of a 512-element K chunk, absmax applied once per (row, block) in fp32).  What changed is everything that kept the
// shared-memory / L1 pipe (LSU) busy besides the one lookup per packed byte, because that pipe -- one 128-byte
// wavefront per clock per SM -- is what bounds this kernel (ncu, profiles/): per 4 KB item
//     round 1:  128 lookups + 32 weight loads + 32 x fragments + ~25 scattered absmax loads / code2 lookups = 217
//     here:     128 lookups + 32 weight loads +  0            +  4                                          = 164
//   * x lives in REGISTERS when a warp always works on the same K chunk (K / 512 divides the warp count: K = 4096,
//     8192 ...).  Only 8 lanes of an MMA feed x (lane (g, t) holds column g of B; the k slots of lane t carry block
//     t or t + 4, so column g is fed where g == t or g == t + 4); those lanes keep their 64-element block (32
//     registers), every other lane keeps zeros.  Phase 0 (blocks 0-3) and phase 1 (blocks 4-7) run into two
//     accumulators: columns 0-3 of the first and 4-7 of the second are the block sums, the other halves collect
//     products of mismatched blocks and are never read -- so B is the same register in both phases and no lane has
//     to mask its x.  Other K: x fragments come from shared memory as in round 1.
//   * absmax is de-nested ONCE per CTA in the prologue (before the dependency wait): coalesced loads of the uint8
//     absmax of the CTA's rows, code2 lookups, fl(fl(code2[q] * absmax2) + offset), fp32 rows in shared memory with a
//     pitch of 8 mod 32 words -- the hot loop reads its four values with two conflict-free 64-bit loads.
//   * weights: ring of ONE item per warp in registers (4 x LDG.256 per item, each half re-issued for the next item
//     as soon as its 16 MMAs are done), plus bulk L2 prefetches of the CTA's whole row range and of the NEXT GEMV's
//     weight (host hint) issued by one warp at kernel entry: HBM -> L2 runs ahead on its own, the ring only has to
//     cover L2 latency.
#pragma once

template <typename T, bool NESTED, int WARPS, bool XREG, int DEPTH, int NACC, bool MULTI>
__global__ void __launch_bounds__(WARPS * 32, WARPS <= 8 ? 2 : 1)
k_gemv4_v2(const GemvArgs a, int x_blocks_padded, int tiles_total, int abs_pitch) {
  // NACC independent accumulator chains per phase: a dependent mma.sync chain costs ~60 cycles per link on this part, and
  // a warp that owns one or two items (the 7B shapes) is bound by exactly that chain of 32 links (phase probe)
  static_assert(NACC == 1 || NACC == 2 || NACC == 4, "accumulator chains");
  static_assert(DEPTH == 1 || (DEPTH == 2 && !XREG), "x in registers leaves room for a ring of one item only");
  // shared memory: [0, 64 KB) byte LUT (entry stride 256 B, one word per lane) | code2 | absmax rows | x | partial sums
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int CT = WARPS * 32;
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("griddepcontrol.launch_dependents;");

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int kb = a.K >> 6;                 // blocks per row
  const int nch = (a.K + 511) >> 9;        // 512-element chunks per row (the last one may be half)
  const int row_bytes = a.K >> 1;
  const int t_begin = (int)(blockIdx.x * (unsigned)tiles_total / gridDim.x);
  const int t_end = (int)((blockIdx.x + 1) * (unsigned)tiles_total / gridDim.x);
  const int ntl = t_end - t_begin;
  const int ntl_max = (tiles_total + (int)gridDim.x - 1) / (int)gridDim.x;

  float *s_code2 = reinterpret_cast<float *>(smem + 65536);
  float *s_abs = reinterpret_cast<float *>(smem + 65536 + 1024);
  unsigned char *s_x = reinterpret_cast<unsigned char *>(s_abs + (size_t)ntl_max * 16 * abs_pitch);
  float *s_part = reinterpret_cast<float *>(s_x + (XREG ? 0 : (size_t)x_blocks_padded * kBcXPitch));   // [tile_local][warp][16]
  const uint32_t x_s = smem_base + (uint32_t)(s_x - smem);

  unsigned long long probe_t = 0, probe_c = 0;
  const bool probing = (a.flags & 2) && blockIdx.x == gridDim.x / 2 && tid == 0;
  if (probing) { probe_t = globaltimer_ns(); probe_c = clock64(); }

  auto mat_of = [&](int tile) { return MULTI ? (int)(tile >= a.mt[1]) + (int)(tile >= a.mt[2]) + (int)(tile >= a.mt[3]) : 0; };
#define BNB_MSEL(arr, m) ((m) == 0 ? a.arr[0] : (m) == 1 ? a.arr[1] : (m) == 2 ? a.arr[2] : a.arr[3])

  // ---- table constants: one round of small loads ahead of everything else
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

  // The table is stored BEFORE the first weight loads are issued: ptxas gives the stores and the loads the same
  // scoreboard slot, so stores placed after the loads wait for the loads' data (phase probe: 0.9 us of every launch).
  // ---- byte LUT: e -> {T(code[e >> 4]), T(code[e & 15])}, replicated for the 32 lanes (bank == lane): 8 threads
  // write one 128-byte entry with conflict-free 128-bit stores
  {
    const int j8 = tid & 7;
    const uint32_t lo16 = MmaT<T>::pack(clo, 0.0f) << 16;
    constexpr int EPI = CT / 8;                      // entries per iteration (32 for 8 warps, 64 for 16)
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
  for (int i = tid; i < ntl * WARPS * 16; i += CT) s_part[i] = 0.f;
  if (probing) g_gemv_probe[9] = globaltimer_ns() - probe_t;        // LUT stored
  // ---- absmax of the CTA's rows: first batch of loads (4 blocks per thread and round), requested AHEAD of the weight stream
  // (behind it they queue for more than a microsecond), de-nested after the dependency wait
  const int total4 = ntl * 4 * kb;                 // groups of 4 consecutive blocks of a row (kb % 4 == 0)
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

  // ---- weights: one 256-bit load per (row, block): a lane owns a whole 32-byte sector, 4 lanes one 128-byte line
  uint32_t w[DEPTH][2][2][8];   // [ring slot][block t / t+4][row half][32 bytes]
  auto load_w = [&](uint32_t (&dst)[2][8], int j, int tile, int c) {
    const int m = mat_of(tile);
    const int lt = MULTI ? tile - BNB_MSEL(mt, m) : tile;
    const int Nm = MULTI ? BNB_MSEL(mN, m) : a.N;
    const unsigned char *Bm = MULTI ? BNB_MSEL(mB, m) : a.B;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int row = min(lt * 16 + g + 8 * h, Nm - 1);
      const unsigned char *p = Bm + (size_t)row * row_bytes + c * 256 + (t + 4 * j) * 32;
      if (c * 8 + t + 4 * j < kb) ld_stream_u8(dst[h], p);
      else {
#pragma unroll
        for (int i = 0; i < 8; i++) dst[h][i] = 0;
      }
    }
  };
  auto advance = [&](int &tl_, int &c_) {
    c_ += WARPS;
    while (c_ >= nch) { c_ -= nch; tl_++; }
  };
  // compute cursor (tl, c) and load cursor (ltl, lc): the register ring keeps DEPTH items per warp in flight
  int tl = 0, c = warp;
  while (c >= nch) { c -= nch; tl++; }
  int ltl = tl, lc = c;
#pragma unroll
  for (int s_ = 0; s_ < DEPTH; s_++) {
    if (ltl < ntl) {
      load_w(w[s_][0], 0, t_begin + ltl, lc);
      load_w(w[s_][1], 1, t_begin + ltl, lc);
    }
    advance(ltl, lc);
  }
  // ---- L2 prefetch, issued by the last warp: the rest of this CTA's rows, then the next GEMV's weight (host hint)
  if (warp == WARPS - 1) {
    if (!MULTI && (a.flags & 4) && ntl > 0) {   // experiment: bulk L2 prefetch of the CTA's own rows
      const size_t r0 = (size_t)t_begin * 16, r1 = min((size_t)t_end * 16, (size_t)a.N);
      const unsigned char *p = a.B + r0 * row_bytes;
      const size_t bytes = (r1 - r0) * row_bytes;
      for (size_t off = (size_t)lane * 8192; off < bytes; off += 32 * 8192) {
        const unsigned int len = (unsigned int)min((size_t)8192, bytes - off) & ~15u;
        if (len) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p + off), "r"(len) : "memory");
      }
    }
#pragma unroll
    for (int u = 0; u < 2; u++)
      if (a.pf_bytes[u]) l2_prefetch_slice(a.pf_ptr[u], a.pf_bytes[u], blockIdx.x, gridDim.x, lane);
  }
  if (probing) g_gemv_probe[7] = globaltimer_ns() - probe_t;        // first loads issued

  if (NESTED && tid < 256) s_code2[tid] = c2v;
  if (probing) g_gemv_probe[2] = globaltimer_ns() - probe_t;        // tables done, about to wait

  asm volatile("griddepcontrol.wait;" ::: "memory");               // x (and out) belong to the previous kernel until here
  if (probing) g_gemv_probe[3] = globaltimer_ns() - probe_t;        // previous kernel complete
  if (a.sig_local != nullptr && a.do_wait) {
    if (tid < a.npeers) {
      const unsigned int target = ld_acquire_sys(a.epoch) * (unsigned int)a.ngroups + (unsigned int)a.gidx;
      const unsigned int *slot = a.sig_local + tid;
      const long long t0 = clock64();
      while ((int)(ld_acquire_sys(slot) - target) < 0) {
        if (clock64() - t0 > 4000000000ll) __trap();
      }
    }
    __syncthreads();
  }

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
      xr[4 * u] = v.x; xr[4 * u + 1] = v.y; xr[4 * u + 2] = v.z; xr[4 * u + 3] = v.w;
    }
  }
  // other K: x goes to shared memory; the first four 16-byte pieces per thread are requested now, stored after the de-nest
  uint4 xv[4];
  const int xpieces = x_blocks_padded * 8, xvalid = a.K >> 3;
  if (!XREG) {
    const uint4 *xg = reinterpret_cast<const uint4 *>(a.x);
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int p = tid + u * CT;
      xv[u] = p < xvalid ? ld_x_u4(xg + p) : make_uint4(0, 0, 0, 0);
    }
  }
  __syncthreads();                                                  // byte LUT, code2 table and zeroed partial sums visible
  // de-nest while the x loads are in flight: the absmax bytes were requested at kernel entry, before the table build
  for (int base = tid + 2 * CT; base < total4; base += 2 * CT) {    // more than 8 blocks per thread: rounds, next loads in flight
    AbsG nx[2];
    abs_load(nx[0], base);
    abs_load(nx[1], base + CT);
    abs_store(ag[0]);
    abs_store(ag[1]);
    ag[0] = nx[0]; ag[1] = nx[1];
  }
  abs_store(ag[0]);
  abs_store(ag[1]);
  if (!XREG) {
    const uint4 *xg = reinterpret_cast<const uint4 *>(a.x);
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int p = tid + u * CT;
      if (p < xpieces) *reinterpret_cast<uint4 *>(s_x + (p >> 3) * kBcXPitch + (p & 7) * 16) = xv[u];
    }
    for (int p0 = tid + 4 * CT; p0 < xpieces; p0 += 4 * CT) {   // long rows: further rounds of four loads per thread
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int p = p0 + u * CT;
        xv[u] = p < xvalid ? ld_x_u4(xg + p) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int p = p0 + u * CT;
        if (p < xpieces) *reinterpret_cast<uint4 *>(s_x + (p >> 3) * kBcXPitch + (p & 7) * 16) = xv[u];
      }
    }
  }
  __syncthreads();                                                  // absmax rows (and x) visible
  if (probing) g_gemv_probe[4] = globaltimer_ns() - probe_t;        // x there

  const uint32_t lane4 = (uint32_t)(lane * 4);
  const uint32_t act0 = (g == t), act1 = (g == t + 4);
  const uint32_t xlane = x_s + g * kBcXPitch;
  uint32_t b0[4] = {0, 0, 0, 0}, b1[4] = {0, 0, 0, 0};
  float acc0 = 0.f, acc1 = 0.f;

  while (tl < ntl) {
#pragma unroll
    for (int s_ = 0; s_ < DEPTH; s_++) {
      if (tl >= ntl) break;
      int ntl_ = tl, nc = c;
      advance(ntl_, nc);
      const bool lhave = ltl < ntl;
      // the four absmax values of this lane's partial sums: rows g, g + 8; blocks 2t, 2t + 1 of the chunk
      const float *ap = s_abs + (tl * 16 + g) * abs_pitch + min(c * 8 + 2 * t, kb - 2);   // half chunk at the end of a row: finite values, zero sums
      const float2 am_lo = *reinterpret_cast<const float2 *>(ap);
      const float2 am_hi = *reinterpret_cast<const float2 *>(ap + 8 * abs_pitch);
      const uint32_t xc = xlane + c * (8 * kBcXPitch);
      float dA[NACC][4], dB[NACC][4];        // phase 0 / phase 1 chains (non-XREG: both phases feed dA)
#pragma unroll
      for (int u = 0; u < NACC; u++) {
#pragma unroll
        for (int v = 0; v < 4; v++) { dA[u][v] = 0.f; dB[u][v] = 0.f; }
      }
#pragma unroll
      for (int j = 0; j < 2; j++) {          // j = 0: blocks 0-3 (columns 0-3), j = 1: blocks 4-7 (columns 4-7)
#pragma unroll
        for (int mg = 0; mg < 8; mg++) {     // one 32-bit word of each row = 8 elements = 2 MMAs
          if (!XREG) {
            if (j == 0) lds_x4_pred(b0, xc + mg * 16, act0);
            else lds_x4_pred(b1, xc + mg * 16, act1);
          }
          const uint32_t s0 = w[s_][j][0][mg], s1 = w[s_][j][1][mg];
#pragma unroll
          for (int mm = 0; mm < 2; mm++) {
            const uint32_t selA = 0x7604u | ((2 * mm) << 4), selB = 0x7604u | ((2 * mm + 1) << 4);
            uint32_t af[4];
            af[0] = *reinterpret_cast<const uint32_t *>(smem + __byte_perm(s0, lane4, selA));
            af[1] = *reinterpret_cast<const uint32_t *>(smem + __byte_perm(s1, lane4, selA));
            af[2] = *reinterpret_cast<const uint32_t *>(smem + __byte_perm(s0, lane4, selB));
            af[3] = *reinterpret_cast<const uint32_t *>(smem + __byte_perm(s1, lane4, selB));
            constexpr int NA = NACC;
            const int ch = (2 * mg + mm) % NA;
            if (XREG) {
              if (j == 0) MmaT<T>::mma(dA[ch], af, xr[(4 * mg + 2 * mm) % (XREG ? 32 : 1)], xr[(4 * mg + 2 * mm + 1) % (XREG ? 32 : 1)]);
              else MmaT<T>::mma(dB[ch], af, xr[(4 * mg + 2 * mm) % (XREG ? 32 : 1)], xr[(4 * mg + 2 * mm + 1) % (XREG ? 32 : 1)]);
            } else {
              if (j == 0) MmaT<T>::mma(dA[ch], af, b0[2 * mm], b0[2 * mm + 1]);
              else MmaT<T>::mma(dA[ch], af, b1[2 * mm], b1[2 * mm + 1]);
            }
          }
        }
        if (lhave) load_w(w[s_][j], j, t_begin + ltl, lc);   // refill this half of the slot: item DEPTH ahead
      }
      advance(ltl, lc);
      float d0[4], d1[4];                    // chains folded in a fixed order
#pragma unroll
      for (int v = 0; v < 4; v++) {
        d0[v] = dA[0][v]; d1[v] = dB[0][v];
#pragma unroll
        for (int u = 1; u < NACC; u++) { d0[v] += dA[u][v]; d1[v] += dB[u][v]; }
      }
      if (XREG) {   // columns 0-3 of d0 and 4-7 of d1 are the block sums (see header)
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
          float *slot = s_part + (tl * WARPS + warp) * 16;
          slot[g] = acc0;
          slot[g + 8] = acc1;
        }
        acc0 = acc1 = 0.f;
      }
      tl = ntl_; c = nc;
    }
  }
  if (probing) g_gemv_probe[5] = globaltimer_ns() - probe_t;        // warp 0 finished its items
  __syncthreads();
  if (probing) g_gemv_probe[6] = globaltimer_ns() - probe_t;        // every warp finished
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
    __syncthreads();
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
  if (probing) { g_gemv_probe[0] = clock64() - probe_c; g_gemv_probe[1] = globaltimer_ns() - probe_t; }
#undef BNB_MSEL
}

// host side: pick the variant, returns false when the shape does not fit (caller falls back to the round-1 kernel)
template <typename T, bool NESTED, bool MULTI>
static bool launch_v2(const GemvArgs &a, int tiles, int sms) {
  static int warps_env = -1, pdl_off = 0, xreg_env = 1, max_rounds = 3, persm_env = 0, nacc_env = 0;
  if (warps_env < 0) {
    const char *e = getenv("BNB_B200_GEMV_V2W"); warps_env = e ? atoi(e) : 0;
    const char *f = getenv("BNB_B200_GEMV_PDL"); pdl_off = (f && f[0] == '0') ? 1 : 0;
    const char *x = getenv("BNB_B200_GEMV_XREG"); xreg_env = x ? atoi(x) : 1;     // 0: never, 1: small shapes, 2: whenever possible
    const char *r = getenv("BNB_B200_GEMV_V2MAX"); max_rounds = r ? atoi(r) : 3;
    const char *p1 = getenv("BNB_B200_GEMV_PERSM"); persm_env = p1 ? atoi(p1) : 0;
    const char *na = getenv("BNB_B200_GEMV_NACC"); nacc_env = na ? atoi(na) : 0;
  }
  const int kb = a.K / 64, nch = ceil_div(a.K, 512);
  const int abs_pitch = ceil_div(kb, 32) * 32 + 8;
  const int xblocks = nch * 8;
  for (int attempt = 0; attempt < 2; attempt++) {
    const int warps = warps_env ? (attempt == 0 ? warps_env : 24 - warps_env) : (attempt == 0 ? 8 : 16);
    if (warps != 8 && warps != 16) continue;
    const int per_sm = (warps == 8 && persm_env != 1) ? 2 : 1;   // PERSM=1: one 8-warp CTA per SM per kernel, the next kernel's CTA co-resident
    const int grid = tiles < sms * per_sm ? tiles : sms * per_sm;
    const int ntl_max = ceil_div(tiles, grid);
    // x in registers costs the second ring slot: taken where a warp has one or two items anyway
    const bool xreg_ok = (a.K % 512 == 0) && (warps % nch == 0);
    const bool xreg = xreg_ok && (xreg_env == 2 || (xreg_env == 1 && ntl_max * nch <= 2 * warps));
    // the absmax of the CTA's rows is de-nested in the prologue: worth it while that is a few rounds per thread (the 7B
    // shapes); a CTA with a long row range keeps the per-item loads of the round-1 kernel (caller falls back)
    if (ntl_max * 4 * kb > max_rounds * warps * 32) continue;
    const size_t need = (size_t)65536 + 1024 + (size_t)ntl_max * 16 * abs_pitch * 4 + (xreg ? 0 : (size_t)xblocks * kBcXPitch) +
                        (size_t)ntl_max * warps * 16 * sizeof(float);
    if (need > (size_t)(warps == 8 ? 113 * 1024 : kBcSmemMax)) continue;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid); lc.blockDim = dim3(warps * 32); lc.dynamicSmemBytes = need; lc.stream = current_stream();
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr; lc.numAttrs = pdl_off ? 0 : 1;
#define V2_LAUNCH(WARPS_, XREG_, DEPTH_, NACC_)                                                                          \
  do {                                                                                                                  \
    auto kfn = k_gemv4_v2<T, NESTED, WARPS_, XREG_, DEPTH_, NACC_, MULTI>;                                              \
    ensure_max_dynamic_smem(reinterpret_cast<const void *>(kfn), kBcSmemMax, "gemv v2 smem attr");                      \
    latch_error(cudaLaunchKernelEx(&lc, kfn, a, xblocks, tiles, abs_pitch), "gemv_4bit (v2) launch");                   \
  } while (0)
    // few items per warp: the accumulator chain is the bound -> ring of one, more chains; many: ring of two, one chain
    const bool few = ntl_max * nch <= 2 * warps;
    if (warps == 8) {
      if (xreg) { if (nacc_env == 1) V2_LAUNCH(8, true, 1, 1); else V2_LAUNCH(8, true, 1, 2); }
      else if (few && nacc_env != 1) V2_LAUNCH(8, false, 1, 4);
      else V2_LAUNCH(8, false, 2, 1);
    } else {
      if (xreg) { if (nacc_env == 1) V2_LAUNCH(16, true, 1, 1); else V2_LAUNCH(16, true, 1, 2); }
      else if (few && nacc_env != 1) V2_LAUNCH(16, false, 1, 4);
      else V2_LAUNCH(16, false, 2, 1);
    }
#undef V2_LAUNCH
    check_launch("gemv_4bit (v2)");
    return true;
  }
  return false;
}
