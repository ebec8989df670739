// The shape phases of one window-sized object, from its 64-bit row masks to ShapeStats (the three chained exact EDTs of
// src/extraction/core/functions/cell.py:176-229), used by object_edt_grid (object_edt.cu).  Include inside the
// translation unit's anonymous namespace after warp_common.cuh.  Shared-memory regions (byte offsets into dyn): rowmask u64[64] (filled by the caller, bit c of
// row r <-> window pixel (r, c), columns may be shifted right by s_lab), topmask u64[64], info u32[64],
// grid u16 [kEdtGridRows][64] (128-byte rows; the caller need not initialise it).
#pragma once

constexpr u32 kMargin = 4;                          // zero rows above and below the window
constexpr u32 kEdtGridRows = kSide + 2 * kMargin;   // 72
constexpr u32 kEdtGridBytes = kEdtGridRows * 128;   // 9 216

// squared distance of column c to the nearest set bit of m (0xFFFFFFFF if m == 0)
__device__ __forceinline__ u32 nearest_bit_sq(u64 m, u32 c) {
  if (m == 0) return kFull;
  const u64 le = m & (~0ull >> (63 - c));  // bits <= c
  const u64 ge = m >> c;                   // bits >= c, shifted
  u32 d = 64;
  if (le) d = c - (63u - (u32)__clzll((long long)le));
  if (ge) d = min(d, (u32)__ffsll((long long)ge) - 1u);
  return d * d;
}

// row distance of an object pixel at column c of row mask m (distance to the nearest zero bit, zeros beyond both ends)
__device__ __forceinline__ u32 row_distance(u64 m, u32 c) {
  const u64 z = ~m;
  const u64 le = c ? (z & (~0ull >> (64 - c))) : 0ull;  // zeros at columns < c
  const u32 dl = le ? (c - (63u - (u32)__clzll((long long)le))) : (c + 1u);
  const u64 ge = c < 63u ? (z >> (c + 1u)) : 0ull;      // zeros at columns > c
  const u32 dr = ge ? (u32)__ffsll((long long)ge) : (64u - c);
  return min(dl, dr);
}

__device__ __forceinline__ void shape_from_masks(const abx_object_rec& rec, u32 s_lab, u32 rowmask_off, u32 topmask_off,
                                                 u32 info_off, u32 g_off, bool zero_top_margin, bool want_conical,
                                                 const double* __restrict__ sqrt_tab, ShapeStats* __restrict__ dst) {
  u64* rowmask = reinterpret_cast<u64*>(dyn + rowmask_off);
  u64* topmask = reinterpret_cast<u64*>(dyn + topmask_off);
  const u32 lane = lane_id();
  const int h = (int)(rec.rmax - rec.rmin) + 1;
  const u32 n = rec.n;
  if (zero_top_margin) *reinterpret_cast<uint4*>(dyn + g_off + 16u * lane) = make_uint4(0, 0, 0, 0);
  topmask[lane] = 0;
  topmask[lane + 32] = 0;
  __syncwarp();
  // ---- phase R: squared row distances, two columns per lane ----
  // lanes over rows first: the run ends of each row, a | b << 8 | several-runs << 16 (an empty row: a = 64, b = 0)
  u32* rowinfo = reinterpret_cast<u32*>(dyn + info_off);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int r = (int)lane + 32 * k;
    if (r < h) {
      const u64 m = rowmask[r];
      u32 info = 64u;
      if (m) {
        const u32 a = (u32)__ffsll((long long)m) - 1u, b = 63u - (u32)__clzll((long long)m);
        const u64 run = m >> a;
        info = a | (b << 8) | (((run & (run + 1ull)) == 0ull) ? 0u : 0x10000u);
      }
      rowinfo[r] = info;
    }
  }
  __syncwarp();
  const u32 c0 = 2u * lane;
#pragma unroll 4
  for (int r = 0; r < h; ++r) {
    const u32 info = rowinfo[r];  // warp-uniform
    const int a = (int)(info & 0xFFu), b = (int)((info >> 8) & 0xFFu);
    // one run [a, b]: min(c - a, b - c) + 1 inside it, <= 0 outside
    u32 g0 = (u32)max(min((int)c0 - a, b - (int)c0) + 1, 0);
    u32 g1 = (u32)max(min((int)c0 + 1 - a, b - (int)c0 - 1) + 1, 0);
    if (info & 0x10000u) {  // several runs (rare)
      const u64 m = rowmask[r];
      g0 = ((m >> c0) & 1ull) ? row_distance(m, c0) : 0u;
      g1 = ((m >> (c0 + 1u)) & 1ull) ? row_distance(m, c0 + 1u) : 0u;
    }
    *reinterpret_cast<u32*>(dyn + g_off + ((u32)r + kMargin) * 128u + 4u * lane) = (g0 * g0) | ((g1 * g1) << 16);
  }
  // the frame row and the margin below the window (the label box may have left other labels there)
  *reinterpret_cast<uint4*>(dyn + g_off + ((u32)h + kMargin) * 128u + 16u * lane) = make_uint4(0, 0, 0, 0);
  __syncwarp();
  // ---- phase C: EDT 1, exact column pass on packed pairs, four target rows per lane at a time ----
  u32 lmax = 0;
  double s_nn = 0.0;
  u64 at_lo = 0, at_hi = 0;  // bit 8 * group + 2 * row-in-group + half: that pixel attains lmax
  {
    u32 grp = 0;
#pragma unroll 1
    for (int r0 = 0; r0 < h; r0 += 4, ++grp) {
      const u32 base = g_off + ((u32)r0 + kMargin) * 128u + 4u * lane;  // byte offset of (row r0, this lane's pair)
      const u32 q0 = *reinterpret_cast<const u32*>(dyn + base);
      const u32 q1 = *reinterpret_cast<const u32*>(dyn + base + 128u);
      const u32 q2 = *reinterpret_cast<const u32*>(dyn + base + 256u);
      const u32 q3 = *reinterpret_cast<const u32*>(dyn + base + 384u);
      constexpr u32 k1 = 0x00010001u, k4 = 0x00040004u, k9 = 0x00090009u;
      // sources inside the group
      u32 b0 = __viaddmin_u16x2(q1, k1, q0); b0 = __viaddmin_u16x2(q2, k4, b0); b0 = __viaddmin_u16x2(q3, k9, b0);
      u32 b1 = __viaddmin_u16x2(q0, k1, q1); b1 = __viaddmin_u16x2(q2, k1, b1); b1 = __viaddmin_u16x2(q3, k4, b1);
      u32 b2 = __viaddmin_u16x2(q0, k4, q2); b2 = __viaddmin_u16x2(q1, k1, b2); b2 = __viaddmin_u16x2(q3, k1, b2);
      u32 b3 = __viaddmin_u16x2(q0, k9, q3); b3 = __viaddmin_u16x2(q1, k4, b3); b3 = __viaddmin_u16x2(q2, k1, b3);
      // sources outside: step d brings row r0 - d (distances d .. d + 3 to the four targets) and row r0 + 3 + d
      u32 e0 = k1, e1 = k4, e2 = k9, e3 = 0x00100010u;  // (d + j)^2 on both halves, d = 1
      u32 inc = 0x00090009u;                             // 2 (d + 4) - 1: e3 of the next step = e3 + inc
      u32 up = base - 128u, dn = base + 512u;
      u32 lim = 1;                                       // d^2: nothing outside is closer than d
      u32 mx = __vmaxu2(__vmaxu2(b0, b1), __vmaxu2(b2, b3));
      mx = max(mx & 0xFFFFu, mx >> 16);
      while (mx > lim) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {  // two steps per test: a surplus step reads one more zero-bounded row
          const u32 ga = *reinterpret_cast<const u32*>(dyn + up);
          const u32 gb = *reinterpret_cast<const u32*>(dyn + dn);
          b0 = __viaddmin_u16x2(ga, e0, b0); b0 = __viaddmin_u16x2(gb, e3, b0);
          b1 = __viaddmin_u16x2(ga, e1, b1); b1 = __viaddmin_u16x2(gb, e2, b1);
          b2 = __viaddmin_u16x2(ga, e2, b2); b2 = __viaddmin_u16x2(gb, e1, b2);
          b3 = __viaddmin_u16x2(ga, e3, b3); b3 = __viaddmin_u16x2(gb, e0, b3);
          e0 = e1; e1 = e2; e2 = e3; e3 += inc; inc += 0x00020002u;
          up -= 128u; dn += 128u;
        }
        lim = e0 & 0xFFFFu;  // (d + 1)^2 of the step that comes next
        mx = __vmaxu2(__vmaxu2(b0, b1), __vmaxu2(b2, b3));
        mx = max(mx & 0xFFFFu, mx >> 16);
      }
      // ---- this lane's eight results: running maximum and who attains it, sum of distances ----
      if (mx > lmax) { lmax = mx; at_lo = at_hi = 0; }
      if (mx == lmax && mx > 0) {
        u32 bits = 0;
        bits |= ((b0 & 0xFFFFu) == lmax) ? 1u : 0u;   bits |= ((b0 >> 16) == lmax) ? 2u : 0u;
        bits |= ((b1 & 0xFFFFu) == lmax) ? 4u : 0u;   bits |= ((b1 >> 16) == lmax) ? 8u : 0u;
        bits |= ((b2 & 0xFFFFu) == lmax) ? 16u : 0u;  bits |= ((b2 >> 16) == lmax) ? 32u : 0u;
        bits |= ((b3 & 0xFFFFu) == lmax) ? 64u : 0u;  bits |= ((b3 >> 16) == lmax) ? 128u : 0u;
        if (grp < 8) at_lo |= (u64)bits << (8u * grp); else at_hi |= (u64)bits << (8u * (grp - 8u));
      }
      if (want_conical) {  // non-object cells have distance 0
        s_nn += sqrt_tab[b0 & 0xFFFFu] + sqrt_tab[b0 >> 16];
        s_nn += sqrt_tab[b1 & 0xFFFFu] + sqrt_tab[b1 >> 16];
        s_nn += sqrt_tab[b2 & 0xFFFFu] + sqrt_tab[b2 >> 16];
        s_nn += sqrt_tab[b3 & 0xFFFFu] + sqrt_tab[b3 >> 16];
      }
    }
  }
  const u32 max_nn2 = __reduce_max_sync(kFull, lmax);
  if (want_conical) {
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) s_nn += __shfl_xor_sync(kFull, s_nn, k);
  }
  // ---- phase T: cone top = pixels with nn2 == max ----
  if (lmax == max_nn2) {
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      u64 m = half ? at_hi : at_lo;
      while (m) {
        const u32 b = (u32)__ffsll((long long)m) - 1u;
        m &= m - 1;
        const u32 row = 4u * ((b >> 3) + 8u * half) + ((b >> 1) & 3u), col = c0 + (b & 1u);
        atomicOr(reinterpret_cast<unsigned long long*>(&topmask[row]), 1ull << col);
      }
    }
  }
  __syncwarp();
  const u64 tm0 = topmask[lane], tm1 = topmask[lane + 32];
  const u32 n_top = __reduce_add_sync(kFull, (u32)(__popcll(tm0) + __popcll(tm1)));
  // ---- phase 2: max over the object of the squared distance to the nearest cone-top pixel ----
  u32 lmax2 = 0;
  if (n_top <= 32) {
    u32 my_top = 0;  // lane k keeps top pixel k (row-major order), (r << 6) | c
    {
      u64 a = tm0, b = tm1;
      u32 k = 0;
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        u64& cur = pass == 0 ? a : b;
        u32 any = __ballot_sync(kFull, cur != 0);
        while (any) {
          const int src = __ffs(any) - 1;
          const u64 mm = __shfl_sync(kFull, cur, src);
          const u32 c = (u32)__ffsll((long long)mm) - 1u;
          const u32 r = (u32)src + 32u * pass;
          if (lane == k) my_top = (r << 6) | c;
          ++k;
          if ((int)lane == src) cur &= cur - 1;
          any = __ballot_sync(kFull, cur != 0);
        }
      }
    }
    if (n_top == 1) {
      // one top pixel: the farthest pixel of a row is one of the row's two ends
      const u32 tp = __shfl_sync(kFull, my_top, 0);
      const int tr = (int)(tp >> 6), tc = (int)(tp & 63u);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int r = (int)lane + 32 * k;
        const u64 m = r < h ? rowmask[r] : 0ull;
        if (m) {
          const int a = __ffsll((long long)m) - 1, b = 63 - __clzll((long long)m);
          const int dc = max(abs(a - tc), abs(b - tc)), dr = r - tr;
          lmax2 = max(lmax2, (u32)(dr * dr + dc * dc));
        }
      }
    } else if (n_top <= 4) {
      // the column terms of the (up to) four tops stay in registers: one packed add-min per (row, top)
      u32 pc[4];
      int tr[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const u32 tp = __shfl_sync(kFull, my_top, (u32)t < n_top ? t : 0);  // unused entries repeat top 0
        tr[t] = (int)(tp >> 6);
        const int d = (int)c0 - (int)(tp & 63u);
        pc[t] = (u32)(d * d) | ((u32)((d + 1) * (d + 1)) << 16);
      }
#pragma unroll 2
      for (int r = 0; r < h; ++r) {
        const u64 m = rowmask[r];
        u32 best = kFull;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int dr = r - tr[t];
          best = __viaddmin_u16x2(pc[t], (u32)(dr * dr) * 0x00010001u, best);
        }
        const u32 pair = (u32)(m >> c0) & 3u;
        if (pair & 1u) lmax2 = max(lmax2, best & 0xFFFFu);
        if (pair & 2u) lmax2 = max(lmax2, best >> 16);
      }
    } else {
      // 5 .. 32 tops: four at a time as above, the per-row minima of the earlier chunks wait in the grid (free by now)
#pragma unroll 1
      for (u32 t0 = 0; t0 < n_top; t0 += 4) {
        u32 pc[4];
        int tr[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const u32 tp = __shfl_sync(kFull, my_top, t0 + (u32)t < n_top ? t0 + (u32)t : t0);  // repeats fill the chunk
          tr[t] = (int)(tp >> 6);
          const int d = (int)c0 - (int)(tp & 63u);
          pc[t] = (u32)(d * d) | ((u32)((d + 1) * (d + 1)) << 16);
        }
        const bool last = t0 + 4u >= n_top;
#pragma unroll 2
        for (int r = 0; r < h; ++r) {
          u32* cell = reinterpret_cast<u32*>(dyn + g_off + ((u32)r + kMargin) * 128u + 4u * lane);
          u32 best = t0 ? *cell : kFull;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int dr = r - tr[t];
            best = __viaddmin_u16x2(pc[t], (u32)(dr * dr) * 0x00010001u, best);
          }
          if (last) {
            const u32 pair = (u32)(rowmask[r] >> c0) & 3u;
            if (pair & 1u) lmax2 = max(lmax2, best & 0xFFFFu);
            if (pair & 2u) lmax2 = max(lmax2, best >> 16);
          } else {
            *cell = best;
          }
        }
      }
    }
  } else {
    // plateau (more than 32 top pixels, e.g. a cell cut straight by the image border): per row that holds top pixels,
    // the squared distance of every column to the row's nearest top pixel — packed pairs, parked in the grid (free by
    // now) — then one packed add-min per (object row, top row)
    u32* toprow = reinterpret_cast<u32*>(dyn + info_off);  // the run ends are no longer needed
    u32 rows0 = __ballot_sync(kFull, tm0 != 0), rows1 = __ballot_sync(kFull, tm1 != 0);
    int nt = 0;
#pragma unroll 1
    while (rows0 | rows1) {
      int r;
      if (rows0) { r = __ffs(rows0) - 1; rows0 &= rows0 - 1; }
      else { r = 32 + __ffs(rows1) - 1; rows1 &= rows1 - 1; }
      const u64 tm = topmask[r];  // warp-uniform
      *reinterpret_cast<u32*>(dyn + g_off + ((u32)nt + kMargin) * 128u + 4u * lane) =
          nearest_bit_sq(tm, c0) | (nearest_bit_sq(tm, c0 + 1u) << 16);
      if (lane == 0) toprow[nt] = (u32)r;
      ++nt;
    }
    __syncwarp();
#pragma unroll 1
    for (int r = 0; r < h; ++r) {
      const u32 pair = (u32)(rowmask[r] >> c0) & 3u;
      u32 best = kFull;
#pragma unroll 2
      for (int j = 0; j < nt; ++j) {
        const int dr = r - (int)toprow[j];
        best = __viaddmin_u16x2(*reinterpret_cast<const u32*>(dyn + g_off + ((u32)j + kMargin) * 128u + 4u * lane),
                                (u32)(dr * dr) * 0x00010001u, best);
      }
      if (pair & 1u) lmax2 = max(lmax2, best & 0xFFFFu);
      if (pair & 2u) lmax2 = max(lmax2, best >> 16);
    }
  }
  const u32 max_dn2 = __reduce_max_sync(kFull, lmax2);
  // ---- phase 3: size of the cone top = distance of each top pixel to the rest of the object ----
  double s_top = 0.0;
  if (n_top == n) {
    // `dn == 0` has no zero at all: SciPy measures to index (-1, 0) of the padded plane
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int r = (int)lane + 32 * k;
      u64 m = r < h ? rowmask[r] : 0ull;
      while (m) {
        const u32 cb = (u32)__ffsll((long long)m) - 1u;
        m &= m - 1;
        const double dr = (double)rec.rmin + (double)r + 2.0, dc = (double)rec.cmin + (double)(cb - s_lab) + 1.0;
        s_top += sqrt(dr * dr + dc * dc);
      }
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) s_top += __shfl_xor_sync(kFull, s_top, k);
  } else {
    // lanes over rows: q = object pixels that are not cone top
    const u64 q0 = (lane < (u32)h) ? (rowmask[lane] & ~tm0) : 0ull;
    const u64 q1 = (lane + 32 < (u32)h) ? (rowmask[lane + 32] & ~tm1) : 0ull;
    u32 rows0 = __ballot_sync(kFull, tm0 != 0), rows1 = __ballot_sync(kFull, tm1 != 0);  // rows that hold top pixels
#pragma unroll 1
    while (rows0 | rows1) {
      int r;
      if (rows0) { r = __ffs(rows0) - 1; rows0 &= rows0 - 1; }
      else { r = 32 + __ffs(rows1) - 1; rows1 &= rows1 - 1; }
      u64 tm = topmask[r];  // warp-uniform
      while (tm) {
        const u32 c = (u32)__ffsll((long long)tm) - 1u;
        tm &= tm - 1;
        u32 best = kFull;
        const u32 d0 = nearest_bit_sq(q0, c);
        const int dr0 = r - (int)lane;
        if (d0 != kFull) best = d0 + (u32)(dr0 * dr0);
        const u32 d1 = nearest_bit_sq(q1, c);
        const int dr1 = r - (int)lane - 32;
        if (d1 != kFull) best = min(best, d1 + (u32)(dr1 * dr1));
        best = __reduce_min_sync(kFull, best);
        s_top += sqrt((double)best);  // same value in every lane
      }
    }
  }
  if (lane == 0) {
    ShapeStats out;
    out.sum_nn = s_nn; out.sum_top = s_top; out.max_nn2 = max_nn2; out.max_dn2 = max_dn2;
    *dst = out;
  }
  __syncwarp();
}

