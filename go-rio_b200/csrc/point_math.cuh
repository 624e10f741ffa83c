// Per-point device functions shared by the streaming kernels (corr.cu,
// linearize.cu: one launch per stage, large clouds) and the device-resident LM
// kernel (lm.cu: the whole optimizer loop of one registration in one launch).
// Both paths run the SAME arithmetic in the SAME order per point.
#pragma once
#include "kernels.cuh"

namespace apd {

constexpr unsigned long long kInfKey = 0xffffffffffffffffull;

// scan one contiguous range of the cell-sorted target points
__device__ __forceinline__ void scan_range(const float4* spts, int b, int e, float qx, float qy, float qz,
                                           unsigned long long& best, int& best_pos) {
  for (int j = b; j < e; j++) {
    const float4 p = spts[j];
    const unsigned long long key = pack_key(sqdist_rn(qx, qy, qz, p.x, p.y, p.z), __float_as_int(p.w));
    if (key < best) {
      best = key;
      best_pos = j;
    }
  }
}

// The same, also keeping `second2`: a lower bound of the squared distance of every scanned point OTHER than the best one
// (the seed of a warm-started search is met again at its own position: it is not "another" point).
template <bool kTrack>
__device__ __forceinline__ void scan_range_t(const float4* spts, int b, int e, float qx, float qy, float qz,
                                             unsigned long long& best, int& best_pos, float& second2) {
  if constexpr (!kTrack) {
    scan_range(spts, b, e, qx, qy, qz, best, best_pos);
  } else {
    for (int j = b; j < e; j++) {
      const float4 p = spts[j];
      const float d = sqdist_rn(qx, qy, qz, p.x, p.y, p.z);
      const unsigned long long key = pack_key(d, __float_as_int(p.w));
      if (key < best) {
        second2 = fminf(second2, __uint_as_float((unsigned)(best >> 32)));  // (no best yet: NaN, which fminf ignores)
        best = key;
        best_pos = j;
      } else if (j != best_pos) {
        second2 = fminf(second2, d);
      }
    }
  }
}

// Exact nearest neighbour by (d2, original index). Expands Chebyshev shells of
// cells until the best distance is provably final, or until every unscanned
// point is farther than `limit` (then the caller rejects the match anyway).
// G consecutive lanes share one query (G = 1, 2, 4, 8, 16, 32): the x-rows of a
// shell are dealt round-robin to the G lanes and the group's best key is
// min-reduced with shuffles after every shell, so the critical path of a query is
// ~1/G of the single-thread scan (small source clouds are latency-bound).
// All G lanes return the same result.
// barrier over the G lanes that share a query (all of them have finished reading what one of them is about to overwrite)
template <int G>
__device__ __forceinline__ void nn_group_sync() {
  if (G > 1) __syncwarp((G >= 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1))));
}

// min-reduce (best, best_pos) over the G lanes that share a query
template <int G>
__device__ __forceinline__ void nn_group_min(unsigned long long& best, int& best_pos) {
  if (G > 1) {
    // groups of one warp leave the shell loop at different times: shuffle within the group's own lanes only
    const unsigned gmask = (G >= 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) {
      const unsigned long long ob = __shfl_xor_sync(gmask, best, o);
      const int op = __shfl_xor_sync(gmask, best_pos, o);
      if (ob < best) {
        best = ob;
        best_pos = op;
      }
    }
  }
}

// The G lanes of a query scan candidate segments TOGETHER: every lane brings one segment [b, b + cnt) of the cell-sorted
// target points (cnt = 0: nothing — its row was pruned or lies outside the grid); the non-empty segments are taken one
// after the other (ballot over the group), broadcast, and scanned with the lanes strided over their points. After a
// warm start one or two rows of a cube survive the pruning: dealing ROWS to lanes (round 1) left ~1 lane in 8 busy in
// the candidate loop (4.2 active threads per instruction, 27 % of the loop kernel's instructions); dealing POINTS does not.
template <int G>
__device__ __forceinline__ unsigned nn_group_mask() {
  return (G >= 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
}
template <int G>
__device__ __forceinline__ void scan_segments_group(const float4* spts, int b, int cnt, float qx, float qy, float qz,
                                                    unsigned long long& best, int& best_pos) {
  if (G == 1) {
    scan_range(spts, b, b + cnt, qx, qy, qz, best, best_pos);
    return;
  }
  const unsigned gmask = nn_group_mask<G>();
  const int gbase = (int)(threadIdx.x & 31) & ~(G - 1), sub = (int)(threadIdx.x & (G - 1));
  unsigned m = __ballot_sync(gmask, cnt > 0) & gmask;
  while (m) {  // (uniform over the group)
    const int src = __ffs(m) - 1;
    m &= m - 1;
    const int bb = __shfl_sync(gmask, b, src), cc = __shfl_sync(gmask, cnt, src);
    for (int j = bb + sub; j < bb + cc; j += G) {
      const float4 p = spts[j];
      const unsigned long long key = pack_key(sqdist_rn(qx, qy, qz, p.x, p.y, p.z), __float_as_int(p.w));
      if (key < best) {
        best = key;
        best_pos = j;
      }
    }
  }
  (void)gbase;
}

// squared distance (with the search's safety factor) from the query to the y/z slab of cell row (y, z)
__device__ __forceinline__ float row_dist2(const GridDesc& g, int y, int z, float qy, float qz) {
  const float mg = 0.002f * g.cell;
  const float loy = g.oy + (float)y * g.cell - mg, hiy = g.oy + (float)(y + 1) * g.cell + mg;
  const float loz = g.oz + (float)z * g.cell - mg, hiz = g.oz + (float)(z + 1) * g.cell + mg;
  const float ddy = fmaxf(0.f, fmaxf(loy - qy, qy - hiy)), ddz = fmaxf(0.f, fmaxf(loz - qz, qz - hiz));
  return (ddy * ddy + ddz * ddz) * 0.9999f;
}

// Shells r = 2, 3, ... around the query's cell (cx,cy,cz), given the best key over the radius-1 cube: continues until
// the best distance is provably final or every unscanned point is farther than the limit.
// proven2: on return every point that was NOT scanned is at least sqrt(min(proven2, best d2)) away.
template <int G>
__device__ __forceinline__ void nn_shells(const float4* spts, const uint32_t* cell_start, const GridDesc& g,
                                          float qx, float qy, float qz, int cx, int cy, int cz, double limit_sq, unsigned long long& best,
                                          int& best_pos, float& proven2) {
  const int sub = (G == 1) ? 0 : (int)(threadIdx.x & (G - 1));
  // Rows inside a shell are skipped when they lie beyond 1.25 x the correspondence distance (not 1 x): the slack is what
  // lets the next pass prove "still unmatched" after a small motion without searching again (warm_start).
  const double prune_sq = limit_sq * 1.5625;
  const float prune2 = prune_sq < 3.0e38 ? (float)prune_sq : 3.4e38f;
  // thick shells (r, rr]: one cell at a time near the query, then growing ~1.5x (see knn_cov.cu)
  for (int r = 1;;) {
    const float lb = ((float)r - 0.002f) * g.cell;
    const float lb2 = lb * lb;
    proven2 = fminf(lb2, prune2);
    if (best != kInfKey && __uint_as_float((unsigned)(best >> 32)) < lb2) break;
    if ((double)lb2 >= limit_sq) break;
    if (cx - r <= 0 && cx + r >= g.nx - 1 && cy - r <= 0 && cy + r >= g.ny - 1 && cz - r <= 0 && cz + r >= g.nz - 1) {
      proven2 = prune2;  // the cube covers the grid; only rows skipped for their distance were left out
      break;
    }
    const int rr = r < 3 ? r + 1 : r + (r >> 1) + 1;
    const int side = 2 * rr + 1;
    const int x0 = max(cx - rr, 0), x1 = min(cx + rr, g.nx - 1);
    // rounds of G rows: lane `sub` sets up row ri0 + sub (its one or two segments), the group scans what survived
    for (int ri0 = 0; ri0 < side * side; ri0 += G) {  // (uniform over the group)
      const int ri = ri0 + sub;
      int b0 = 0, c0 = 0, b1 = 0, c1 = 0;
      if (ri < side * side) {
        const int dy = ri % side - rr, dz = ri / side - rr;
        const int y = cy + dy, z = cz + dz;
        if (y >= 0 && y < g.ny && z >= 0 && z < g.nz) {
          // skip the row if even its nearest point cannot beat the current best / the limit
          const float dyz2 = row_dist2(g, y, z, qy, qz);
          if ((double)dyz2 < prune_sq && !(best != kInfKey && dyz2 > __uint_as_float((unsigned)(best >> 32)))) {
            const int row = (z * g.ny + y) * g.nx;
            if (dy > r || dy < -r || dz > r || dz < -r) {  // row outside the scanned cube: its whole x-range
              b0 = (int)cell_start[row + x0];
              c0 = (int)cell_start[row + x1 + 1] - b0;
            } else {  // row crosses the scanned cube: the two end pieces
              const int xl = min(cx - r - 1, g.nx - 1), xr = max(cx + r + 1, 0);
              if (x0 <= xl) {
                b0 = (int)cell_start[row + x0];
                c0 = (int)cell_start[row + xl + 1] - b0;
              }
              if (xr <= x1) {
                b1 = (int)cell_start[row + xr];
                c1 = (int)cell_start[row + x1 + 1] - b1;
              }
            }
          }
        }
      }
      scan_segments_group<G>(spts, b0, c0, qx, qy, qz, best, best_pos);
      scan_segments_group<G>(spts, b1, c1, qx, qy, qz, best, best_pos);
    }
    nn_group_min<G>(best, best_pos);
    r = rr;
  }
}

// ---- one lane per query (the streaming kernels of large clouds, where the searches are bound by instruction issue):
// the same exact search with the set-up arithmetic taken out of the row loops. The squared distances from the query to
// the (slightly grown) cell slabs at offsets -1 / 0 / +1 are computed once per axis — they are the terms of row_dist2 —
// so that a row of the 3x3x3 cube costs one add and one compare, its end cells are dropped when even their corner is
// farther than the best so far, and in the shells a whole z-slab of rows goes with one test.
__device__ __forceinline__ float slab_dist2(float q, float o, int c, float cell, float mg) {
  const float lo = o + (float)c * cell - mg, hi = o + (float)(c + 1) * cell + mg;
  const float d = fmaxf(0.f, fmaxf(lo - q, q - hi));
  return d * d;
}
// kTrack: also returns second2, a lower bound of the squared distance of EVERY target point other than the one returned —
// the scanned ones by their distances, the rows / cells / slabs that were skipped by the bound they were skipped with, the
// rest by proven2. It is what lets the next pass over the same clouds keep a match without searching (corr.cu).
template <bool kTrack>
__device__ __forceinline__ void nn_search_lane_t(const float4* spts, const uint32_t* cell_start, const GridDesc& g, float qx, float qy, float qz,
                                                 double limit_sq, unsigned long long& best, int& best_pos, int seed_pos, float& proven2,
                                                 float& second2) {
  second2 = 3.4e38f;
  const int cx = cell_coord(qx, g.ox, g.inv_cell, g.nx);
  const int cy = cell_coord(qy, g.oy, g.inv_cell, g.ny);
  const int cz = cell_coord(qz, g.oz, g.inv_cell, g.nz);
  best = kInfKey;
  best_pos = -1;
  if (seed_pos >= 0) {
    const float4 p = spts[seed_pos];
    best = pack_key(sqdist_rn(qx, qy, qz, p.x, p.y, p.z), __float_as_int(p.w));
    best_pos = seed_pos;
  }
  const float mg = 0.002f * g.cell;
  float dx2[3], dy2[3], dz2[3];
#pragma unroll
  for (int o = 0; o < 3; o++) {
    dx2[o] = slab_dist2(qx, g.ox, cx + o - 1, g.cell, mg);
    dy2[o] = slab_dist2(qy, g.oy, cy + o - 1, g.cell, mg);
    dz2[o] = slab_dist2(qz, g.oz, cz + o - 1, g.cell, mg);
  }
  // ring 0+1: the query's own row first; then the rows among the other eight that can still hold a closer point. Which
  // ones is decided for all eight at once (straight-line code), and the lanes of a warp then walk their own survivors
  // side by side — a lane that keeps two rows and a lane that keeps five share two turns, not nine.
  auto scan_row = [&](int oy, int oz, float ryz, bool own) {
    const bool have = best != kInfKey;
    const float bd = __uint_as_float((unsigned)(best >> 32));
    if (have && !own && ryz * 0.9999f > bd) {  // even the row's nearest point cannot beat the current best
      if (kTrack) second2 = fminf(second2, ryz * 0.9999f);
      return;
    }
    // the end cells of the row are only read if their corner can
    const float ea = (ryz + dx2[0]) * 0.9999f, eb = (ryz + dx2[2]) * 0.9999f;
    const bool skip_a = have && ea > bd, skip_b = have && eb > bd;
    const int xa = (cx > 0 && !skip_a) ? cx - 1 : cx;
    const int xb = (cx < g.nx - 1 && !skip_b) ? cx + 1 : cx;
    if (kTrack) {
      if (cx > 0 && skip_a) second2 = fminf(second2, ea);
      if (cx < g.nx - 1 && skip_b) second2 = fminf(second2, eb);
    }
    const int row = ((cz + oz - 1) * g.ny + (cy + oy - 1)) * g.nx;
    scan_range_t<kTrack>(spts, (int)cell_start[row + xa], (int)cell_start[row + xb + 1], qx, qy, qz, best, best_pos, second2);
  };
  if (cy >= 0 && cy < g.ny && cz >= 0 && cz < g.nz) scan_row(1, 1, dy2[1] + dz2[1], true);
  unsigned m = 0;
  {
    const bool have = best != kInfKey;
    const float bd = __uint_as_float((unsigned)(best >> 32));
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int idx = k < 4 ? k : k + 1, oy = idx % 3, oz = idx / 3;  // (compile-time)
      const int y = cy + oy - 1, z = cz + oz - 1;
      const bool inside = y >= 0 && y < g.ny && z >= 0 && z < g.nz;
      const float rb = (dy2[oy] + dz2[oz]) * 0.9999f;
      const bool far = have && rb > bd;
      if (kTrack && inside && far) second2 = fminf(second2, rb);
      m |= (inside && !far) ? (1u << k) : 0u;
    }
  }
  while (m) {
    const int k = __ffs(m) - 1;
    m &= m - 1;
    const int idx = k < 4 ? k : k + 1;
    const int oz = (idx * 11) >> 5, oy = idx - 3 * oz;  // idx / 3, idx % 3 for idx < 9
    const float ry = oy == 0 ? dy2[0] : (oy == 1 ? dy2[1] : dy2[2]);
    const float rz = oz == 0 ? dz2[0] : (oz == 1 ? dz2[1] : dz2[2]);
    scan_row(oy, oz, ry + rz, false);
  }
  // shells r = 2, 3, ... (see nn_shells): z-slabs outermost, their distance hoisted
  const double prune_sq = limit_sq * 1.5625;
  const float prune2 = prune_sq < 3.0e38 ? (float)prune_sq : 3.4e38f;
  for (int r = 1;;) {
    const float lb = ((float)r - 0.002f) * g.cell;
    const float lb2 = lb * lb;
    proven2 = fminf(lb2, prune2);
    if (best != kInfKey && __uint_as_float((unsigned)(best >> 32)) < lb2) break;
    if ((double)lb2 >= limit_sq) break;
    if (cx - r <= 0 && cx + r >= g.nx - 1 && cy - r <= 0 && cy + r >= g.ny - 1 && cz - r <= 0 && cz + r >= g.nz - 1) {
      proven2 = prune2;  // the cube covers the grid; only rows skipped for their distance were left out
      break;
    }
    const int rr = r < 3 ? r + 1 : r + (r >> 1) + 1;
    const int x0 = max(cx - rr, 0), x1 = min(cx + rr, g.nx - 1);
    const int xl = min(cx - r - 1, g.nx - 1), xr = max(cx + r + 1, 0);
    for (int z = max(cz - rr, 0); z <= min(cz + rr, g.nz - 1); z++) {
      const float sz2 = slab_dist2(qz, g.oz, z, g.cell, mg);
      if ((double)(sz2 * 0.9999f) >= prune_sq) continue;
      if (best != kInfKey && sz2 * 0.9999f > __uint_as_float((unsigned)(best >> 32))) {  // the whole slab of rows
        if (kTrack) second2 = fminf(second2, sz2 * 0.9999f);
        continue;
      }
      const bool z_outer = z > cz + r || z < cz - r;
      for (int y = max(cy - rr, 0); y <= min(cy + rr, g.ny - 1); y++) {
        const float dyz2 = (slab_dist2(qy, g.oy, y, g.cell, mg) + sz2) * 0.9999f;
        if ((double)dyz2 >= prune_sq) continue;
        if (best != kInfKey && dyz2 > __uint_as_float((unsigned)(best >> 32))) {
          if (kTrack) second2 = fminf(second2, dyz2);
          continue;
        }
        const int row = (z * g.ny + y) * g.nx;
        if (z_outer || y > cy + r || y < cy - r) {  // row outside the scanned cube: its whole x-range
          scan_range_t<kTrack>(spts, (int)cell_start[row + x0], (int)cell_start[row + x1 + 1], qx, qy, qz, best, best_pos, second2);
        } else {  // row crosses the scanned cube: the two end pieces
          if (x0 <= xl) scan_range_t<kTrack>(spts, (int)cell_start[row + x0], (int)cell_start[row + xl + 1], qx, qy, qz, best, best_pos, second2);
          if (xr <= x1) scan_range_t<kTrack>(spts, (int)cell_start[row + xr], (int)cell_start[row + x1 + 1], qx, qy, qz, best, best_pos, second2);
        }
      }
    }
    r = rr;
  }
  if (kTrack) second2 = fminf(second2, proven2);  // everything that was never looked at
}
__device__ __forceinline__ void nn_search_lane(const float4* spts, const uint32_t* cell_start, const GridDesc& g, float qx, float qy, float qz,
                                               double limit_sq, unsigned long long& best, int& best_pos, int seed_pos, float& proven2) {
  float second2;
  nn_search_lane_t<false>(spts, cell_start, g, qx, qy, qz, limit_sq, best, best_pos, seed_pos, proven2, second2);
}

// seed_pos >= 0: a target point (sorted position) to start from — the previous iteration's match. Its distance
// bounds the search: rows of the cube whose box is farther are skipped, and the shells usually end at once.
// The result is the same exact nearest neighbour by (d2, original index) with or without the seed.
template <int G>
__device__ __forceinline__ void nn_search(const float4* spts, const uint32_t* cell_start, const GridDesc& g,
                                          float qx, float qy, float qz, double limit_sq, unsigned long long& best, int& best_pos,
                                          int seed_pos, float& proven2) {
  if (G == 1) {
    nn_search_lane(spts, cell_start, g, qx, qy, qz, limit_sq, best, best_pos, seed_pos, proven2);
    return;
  }
  const int sub = (G == 1) ? 0 : (int)(threadIdx.x & (G - 1));
  const int cx = cell_coord(qx, g.ox, g.inv_cell, g.nx);
  const int cy = cell_coord(qy, g.oy, g.inv_cell, g.ny);
  const int cz = cell_coord(qz, g.oz, g.inv_cell, g.nz);
  best = kInfKey;
  best_pos = -1;
  if (seed_pos >= 0) {
    const float4 p = spts[seed_pos];
    best = pack_key(sqdist_rn(qx, qy, qz, p.x, p.y, p.z), __float_as_int(p.w));
    best_pos = seed_pos;
  }
  // ring 0+1: the 3x3x3 cube as 9 x-rows. The query's own row first, by all lanes; its best prunes the other eight,
  // which are set up one per lane (G = 8; fewer lanes: in rounds) and scanned together.
  {
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.nx - 1);
    {
      const int row = (cz * g.ny + cy) * g.nx;
      const int b = (int)cell_start[row + x0];
      scan_segments_group<G>(spts, sub == 0 ? b : 0, sub == 0 ? (int)cell_start[row + x1 + 1] - b : 0, qx, qy, qz, best, best_pos);
      nn_group_min<G>(best, best_pos);
    }
    for (int k0 = 0; k0 < 8; k0 += G) {  // (uniform over the group)
      const int kk = k0 + sub;          // the eight rows around the centre one
      int b = 0, cnt = 0;
      if (kk < 8) {
        const int ri = kk < 4 ? kk : kk + 1;
        const int y = cy + (ri % 3) - 1, z = cz + (ri / 3) - 1;
        if (y >= 0 && y < g.ny && z >= 0 && z < g.nz &&
            !(best != kInfKey && row_dist2(g, y, z, qy, qz) > __uint_as_float((unsigned)(best >> 32)))) {
          const int row = (z * g.ny + y) * g.nx;
          b = (int)cell_start[row + x0];
          cnt = (int)cell_start[row + x1 + 1] - b;
        }
      }
      scan_segments_group<G>(spts, b, cnt, qx, qy, qz, best, best_pos);
    }
    nn_group_min<G>(best, best_pos);
  }
  nn_shells<G>(spts, cell_start, g, qx, qy, qz, cx, cy, cz, limit_sq, best, best_pos, proven2);
}
template <int G>
__device__ __forceinline__ void nn_search(const float4* spts, const uint32_t* cell_start, const GridDesc& g,
                                          float qx, float qy, float qz, double limit_sq, unsigned long long& best, int& best_pos) {
  float proven2;
  nn_search<G>(spts, cell_start, g, qx, qy, qz, limit_sq, best, best_pos, -1, proven2);
}

// Warm start of update_correspondences from the previous pass over the same clouds (pose T_prev): decides what the
// search of source point `a` needs. Returns false when the point is PROVABLY still unmatched (no search needed):
// it was unmatched with every target point at least sqrt(prev_sqd) away, it has moved by delta since, and
// sqrt(prev_sqd) - delta still exceeds the correspondence distance (with a 1e-3 margin over the fp32 rounding of d2).
// Otherwise seed_pos is the previous match (or -1) for nn_search.
__device__ __forceinline__ bool warm_start(const PoseF& T_prev, const float4& a, float px, float py, float pz, int prev_corr, float prev_sqd,
                                           double thr_sq, int& seed_pos, float& kept_sqd) {
  seed_pos = prev_corr >= 0 ? (prev_corr & kCorrIndexMask) : -1;
  if (prev_corr >= 0) return true;
  float ox, oy, oz;
  transform_rn(T_prev, a.x, a.y, a.z, ox, oy, oz);
  const float delta = sqrtf(sqdist_rn(px, py, pz, ox, oy, oz)) * 1.0001f;
  const float lb = sqrtf(prev_sqd) * 0.9999f - delta;  // every target point is at least this far from the moved point
  const float thr = (float)sqrt(thr_sq) * 1.001f;
  if (!(prev_sqd < 3.0e38f) || !(lb > thr)) return true;  // no usable bound, or too close to call: search
  kept_sqd = lb * lb;
  return false;
}

__device__ __forceinline__ PoseF pose_to_f32(const PoseD& T) {
  PoseF f;
#pragma unroll
  for (int i = 0; i < 9; i++) f.r[i] = (float)T.r[i];  // Isometry3d::cast<float>() (:164)
#pragma unroll
  for (int i = 0; i < 3; i++) f.t[i] = (float)T.t[i];
  return f;
}


// Radar noise covariance at the transformed source point (px,py,pz) (reference
// fast_apdgicp_impl.hpp:194-210), combined covariance
// RCR = (C_B + C_r) + R (C_A + C_r) R^T (:213-215) and its inverse, the per-point
// Mahalanobis matrix (:217-218). ca_in / cb_in: the regularised covariances of the
// source point and of its matched target point (symmetric-6).
__device__ __forceinline__ Sym3 mahalanobis_of(float px, float py, float pz, const double* ca_in,
                                               const double* cb_in, const PoseD& T, const NoiseParams& np) {
  Sym3 cr;  // cov_r (:204-210); FastGICP has none: adding exact zeros below leaves C_A and C_B as they are
#pragma unroll
  for (int e = 0; e < 6; e++) cr.v[e] = 0.0;
  if (!np.gicp) {
    // radar noise covariance at the transformed point (:194-210)
    const double dpx = (double)px, dpy = (double)py, dpz = (double)pz;
    const double dist = sqrt(dpx * dpx + dpy * dpy + dpz * dpz);
    const double s_x = dist * np.dist_var / 400;
    const double s_y = dist * np.sin_az;
    const double s_z = dist * np.sin_el;
    const float rho_xy = __fsqrt_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)));
    // float-valued angles as in the reference (atan2f); evaluated in double and rounded to float
    const double elevation = (double)(float)atan2((double)rho_xy, dpz);
    const double azimuth = (double)(float)atan2(dpy, dpx);
    double sz_, cz_, sy_, cy_;
    sincos(azimuth * 0.5, &sz_, &cz_);
    sincos(elevation * 0.5, &sy_, &cy_);
    // quaternion of AngleAxis(az, Z) * AngleAxis(el, Y) -> rotation matrix (Eigen toRotationMatrix)
    const double qw = cz_ * cy_, qx = -(sz_ * sy_), qy = cz_ * sy_, qz = sz_ * cy_;
    const double tx = 2.0 * qx, ty = 2.0 * qy, tz = 2.0 * qz;
    const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
    const double txx = tx * qx, txy = ty * qx, txz = tz * qx;
    const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
    double R[9];
    R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.0 - (txx + tyy);
    const double sc[3] = {s_x, s_y, s_z};
    double A[9];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) A[r * 3 + c] = R[r * 3 + c] * sc[c];
    // cov_r = A A^T
    cr.v[0] = A[0] * A[0] + A[1] * A[1] + A[2] * A[2];
    cr.v[1] = A[0] * A[3] + A[1] * A[4] + A[2] * A[5];
    cr.v[2] = A[0] * A[6] + A[1] * A[7] + A[2] * A[8];
    cr.v[3] = A[3] * A[3] + A[4] * A[4] + A[5] * A[5];
    cr.v[4] = A[3] * A[6] + A[4] * A[7] + A[5] * A[8];
    cr.v[5] = A[6] * A[6] + A[7] * A[7] + A[8] * A[8];
  }

  // RCR = (cov_B + cov_r) + T (cov_A + cov_r) T^T (:213-215), 3x3 block
  Sym3 ca, cb;
#pragma unroll
  for (int e = 0; e < 6; e++) {
    ca.v[e] = ca_in[e] + cr.v[e];
    cb.v[e] = cb_in[e] + cr.v[e];
  }
  // X = Rt * ca (3x3 full), then RCR = cb + X * Rt^T
  const double* Rt = T.r;
  const double cam[9] = {ca.v[0], ca.v[1], ca.v[2], ca.v[1], ca.v[3], ca.v[4], ca.v[2], ca.v[4], ca.v[5]};
  double X[9];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) X[r * 3 + c] = Rt[r * 3 + 0] * cam[0 * 3 + c] + Rt[r * 3 + 1] * cam[1 * 3 + c] + Rt[r * 3 + 2] * cam[2 * 3 + c];
  Sym3 rcr;
  const int RR[6] = {0, 0, 0, 1, 1, 2}, CC[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
  for (int e = 0; e < 6; e++) {
    const int r = RR[e], c = CC[e];
    rcr.v[e] = cb.v[e] + (X[r * 3 + 0] * Rt[c * 3 + 0] + X[r * 3 + 1] * Rt[c * 3 + 1] + X[r * 3 + 2] * Rt[c * 3 + 2]);
  }
  return sym_inverse(rcr);  // :217-218
}

// One term of the linearize / compute_error sums (reference :255-293, :318-341):
// e = b - T a, weighted Mahalanobis cost, H_i = J^T M J and b_i = J^T M e with
// J = [skew(T a), -I]. acc: 28 values (21 upper-triangular H, 6 b, err at [27])
// when kHB, else 1 value (err at [0]). `valid` = the point has a correspondence.
template <bool kHB>
__device__ __forceinline__ void accumulate_point(double* acc, bool valid, const float4& a, const float4& b, const double* m_in, double geo_in,
                                                 int c, const PoseD& T, double cl_w) {
  double m[6];
#pragma unroll
  for (int e = 0; e < 6; e++) m[e] = valid ? m_in[e] : 0.0;
  // (invalid points may hold padding bytes: zero them so 0 * garbage cannot make a NaN)
  const double ax = valid ? (double)a.x : 0.0, ay = valid ? (double)a.y : 0.0, az = valid ? (double)a.z : 0.0;
  const double bx = valid ? (double)b.x : 0.0, by = valid ? (double)b.y : 0.0, bz = valid ? (double)b.z : 0.0;
  // transed_mean_A = T * mean_A ; error = mean_B - transed_mean_A (:262-263)
  const double x = ((T.r[0] * ax + T.r[1] * ay) + T.r[2] * az) + T.t[0];
  const double y = ((T.r[3] * ax + T.r[4] * ay) + T.r[5] * az) + T.t[1];
  const double z = ((T.r[6] * ax + T.r[7] * ay) + T.r[8] * az) + T.t[2];
  const double e0 = bx - x, e1 = by - y, e2 = bz - z;
  // M e and the weighted cost (:276)
  const double me0 = (m[0] * e0 + m[1] * e1) + m[2] * e2;
  const double me1 = (m[1] * e0 + m[3] * e1) + m[4] * e2;
  const double me2 = (m[2] * e0 + m[4] * e1) + m[5] * e2;
  const double q = (e0 * me0 + e1 * me1) + e2 * me2;
  const double w = (1.0 + (valid ? geo_in : 0.0)) + ((c & kCorrLabelBit) ? cl_w : 0.0);
  if (!kHB) {
    acc[0] += valid ? w * q : 0.0;
  } else {
    acc[27] += valid ? w * q : 0.0;
    // N = M * skew(t), t = (x,y,z): N[:,0] = M[:,1] z - M[:,2] y, N[:,1] = M[:,2] x - M[:,0] z, N[:,2] = M[:,0] y - M[:,1] x
    const double n00 = m[1] * z - m[2] * y, n01 = m[2] * x - m[0] * z, n02 = m[0] * y - m[1] * x;
    const double n10 = m[3] * z - m[4] * y, n11 = m[4] * x - m[1] * z, n12 = m[1] * y - m[3] * x;
    const double n20 = m[4] * z - m[5] * y, n21 = m[5] * x - m[2] * z, n22 = m[2] * y - m[4] * x;
    // top-left S^T M S (symmetric): TL[r][c] = sum_k S[k][r] N[k][c]
    acc[0] += z * n10 - y * n20;    // (0,0)
    acc[1] += z * n11 - y * n21;    // (0,1)
    acc[2] += z * n12 - y * n22;    // (0,2)
    acc[6] += x * n21 - z * n01;    // (1,1)
    acc[7] += x * n22 - z * n02;    // (1,2)
    acc[11] += y * n02 - x * n12;   // (2,2)
    // top-right -S^T M = -N^T: H(r, 3+c) = -N[c][r]
    acc[3] -= n00;  acc[4] -= n10;  acc[5] -= n20;    // row 0
    acc[8] -= n01;  acc[9] -= n11;  acc[10] -= n21;   // row 1
    acc[12] -= n02; acc[13] -= n12; acc[14] -= n22;   // row 2
    // bottom-right M
    acc[15] += m[0]; acc[16] += m[1]; acc[17] += m[2];
    acc[18] += m[3]; acc[19] += m[4];
    acc[20] += m[5];
    // b = J^T M e = [S^T (M e); -(M e)]
    acc[21] += z * me1 - y * me2;
    acc[22] += x * me2 - z * me0;
    acc[23] += y * me0 - x * me1;
    acc[24] -= me0; acc[25] -= me1; acc[26] -= me2;
  }
}

}  // namespace apd
