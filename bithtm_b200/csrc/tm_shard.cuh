// Segment-sharded temporal memory: the exchange record of a rank and its merge
// (include/bithtm_b200.h, "segment-sharded temporal memory").
//
// The synapse rows are dealt to ranks by segment id; everything the reference orders
// GLOBALLY by segment id -- the matching list (= index of the jitter draw,
// projections.py:235, 247), the learning list (= row of the priority matrix, :120,
// :281) and the recycling choice (lowest ids first, :80-81) -- is rebuilt on every rank
// from the all-gathered per-rank records, which are ascending in id, by counting for
// each entry how many entries of the other ranks precede it (binary searches).
#pragma once

#include "tm_kernels.cuh"

struct XchRecord {  // views into one rank's record of bh_tm_shard_xch_ints() int32
  int* hdr;         // [4]: matching count, recyclable sent, recyclable true count, status
  int* id;          // [xm_cap] matching segment ids, ascending
  int* pot;         // [xm_cap] segment_potential
  int* conn;        // [xm_cap] connected-active count
  int* recyc;       // [xr_cap] lowest recyclable segment ids, ascending
};

__host__ __device__ __forceinline__ long long xch_ints(const bh_ctx& c) { return 4 + 3LL * c.xm_cap + c.xr_cap; }

__device__ __forceinline__ XchRecord xch_record(const bh_ctx& c, int* base, int rank, long long stride = 0) {
  XchRecord r;
  r.hdr = base + rank * (stride ? stride : xch_ints(c));
  r.id = r.hdr + 4;
  r.pot = r.id + c.xm_cap;
  r.conn = r.pot + c.xm_cap;
  r.recyc = r.conn + c.xm_cap;
  return r;
}

// After ph_activate_a (which left the per-CTA counts of matching and recyclable local
// rows in BLK_MATCH / BLK_RECYC): ordered compaction of this rank's record.
__device__ void ph_shard_pack(const bh_ctx& c, int* send, int b, int nb) {
  __shared__ int s_red[32];
  const int NT = blockDim.x;
  const XchRecord rec = xch_record(c, send, 0);
  const int S = c.sc[BH_SC_NSEG];
  const int thr = c.seg_matching_threshold;
  int m_before, m_total, r_before, r_total;
  blk_prefix(BLK(c, BLK_MATCH), b, nb, s_red, m_before, m_total);
  blk_prefix(BLK(c, BLK_RECYC), b, nb, s_red, r_before, r_total);
  const Range rg = block_range(seg_local_count(c, S), b, nb);
  int mbase = m_before, rbase = r_before;
#pragma unroll 1
  for (int tile = rg.begin; tile < rg.end; tile += NT) {
    const int row = tile + threadIdx.x;
    const bool ok = row < rg.end;
    const int s = ok ? seg_gid(c, row) : 0;
    const int pot = ok ? c.seg_pot[s] : 0;
    const bool match = ok && pot >= thr;
    const bool rec_ok = ok && c.seg_count[s] < thr;
    int tot;
    int pos = mbase + block_excl_scan(match ? 1 : 0, s_red, tot);
    mbase += tot;
    if (match && pos < c.xm_cap) {
      rec.id[pos] = s;
      rec.pot[pos] = pot;
      rec.conn[pos] = c.seg_conn[s];
    }
    pos = rbase + block_excl_scan(rec_ok ? 1 : 0, s_red, tot);
    rbase += tot;
    if (rec_ok && pos < c.xr_cap) rec.recyc[pos] = s;
  }
  if (b == 0 && threadIdx.x == 0) {
    // every rank learns of any rank's overflow / tie note through the record (one exchange later)
    int st = c.sc[BH_SC_STATUS];
    if (m_total > c.xm_cap) st |= BH_ST_XCH_OVERFLOW;
    rec.hdr[0] = m_total < c.xm_cap ? m_total : c.xm_cap;
    rec.hdr[1] = r_total < c.xr_cap ? r_total : c.xr_cap;
    rec.hdr[2] = r_total;
    rec.hdr[3] = st;
  }
}

// entries of list[0..n) (ascending) smaller than key.  Gathered records may have been written
// by a peer GPU (shard_fused.cuh): they are read with ld.cv, never through L1.
__device__ __forceinline__ int lower_count(const int* list, int n, int key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldcv(list + mid) < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// Merge the seg_world gathered records (rank order) into the global ascending lists:
// m_seg / m_conn (+ seg_pot / seg_conn of those segments) and recyc_list.
__device__ void ph_shard_merge(const bh_ctx& c, const int* recv, int b, int nb, long long stride = 0) {
  int* base = const_cast<int*>(recv);
  const int G = c.seg_world;
  const int gid0 = b * blockDim.x + threadIdx.x, gsz = nb * blockDim.x;
#pragma unroll 1
  for (long long idx = gid0; idx < (long long)G * c.xm_cap; idx += gsz) {
    const int g = (int)(idx / c.xm_cap), j = (int)(idx - (long long)g * c.xm_cap);
    const XchRecord me = xch_record(c, base, g, stride);
    if (j >= __ldcv(me.hdr)) continue;
    const int s = __ldcv(me.id + j);
    int pos = j;
    for (int o = 0; o < G; ++o)
      if (o != g) {
        const XchRecord other = xch_record(c, base, o, stride);
        pos += lower_count(other.id, __ldcv(other.hdr), s);
      }
    const int pot = __ldcv(me.pot + j), conn = __ldcv(me.conn + j);
    c.seg_pot[s] = pot;  // every rank knows the potentials of the matching segments
    c.seg_conn[s] = conn;
    if (pos < c.match_capacity) {
      c.m_seg[pos] = s;
      c.m_conn[pos] = conn;
    }
  }
#pragma unroll 1
  for (long long idx = gid0; idx < (long long)G * c.xr_cap; idx += gsz) {
    const int g = (int)(idx / c.xr_cap), j = (int)(idx - (long long)g * c.xr_cap);
    const XchRecord me = xch_record(c, base, g, stride);
    if (j >= __ldcv(me.hdr + 1)) continue;
    const int s = __ldcv(me.recyc + j);
    int pos = j;
    for (int o = 0; o < G; ++o)
      if (o != g) {
        const XchRecord other = xch_record(c, base, o, stride);
        pos += lower_count(other.recyc, __ldcv(other.hdr + 1), s);
      }
    c.recyc_list[pos] = s;
  }
  if (gid0 == 0) {
    int m = 0, ra = 0, rt = 0, st = 0;
    for (int g = 0; g < G; ++g) {
      const XchRecord r = xch_record(c, base, g, stride);
      m += __ldcv(r.hdr);
      ra += __ldcv(r.hdr + 1);
      rt += __ldcv(r.hdr + 2);
      st |= __ldcv(r.hdr + 3);
    }
    c.sc[BH_SC_X_MATCH] = m;
    c.sc[BH_SC_X_RECYC_AVAIL] = ra;
    c.sc[BH_SC_X_RECYC_TOTAL] = rt;
    if (st) atomicOr(&c.sc[BH_SC_STATUS], st);
  }
}

// After the merge and draw #3: jittered potential by rank in the global matching list
// (projections.py:234-235), per-cell maximum (:236-237), active-segment count (:251).
// Every rank computes all of it (the per-cell state is replicated).  Completes the step.
// `ready`: draw #3 was bookkeeping only and may still be running on the drawing CTA.
__device__ void ph_activate_finish(const bh_ctx& c, int b, int nb, bool ready = false, bool want_jitter = true) {
  const int mx = c.sc[BH_SC_X_MATCH] < c.match_capacity ? c.sc[BH_SC_X_MATCH] : c.match_capacity;
  const int M = (ready || !want_jitter) ? mx : c.sc[BH_SC_M];
  const long long off3 = c.rng64[R_OFF3], n3 = ready ? (long long)mx : c.rng64[R_N3];
#pragma unroll 1
  for (int j = b * blockDim.x + threadIdx.x; j < M; j += nb * blockDim.x) {
    const int s = c.m_seg[j];
    const int conn = c.m_conn[j];
    const int owner = c.seg_owner[s];
    if (want_jitter) {
      const int pot = c.seg_pot[s];
      const double u = j < n3 ? rng_uniform(c, off3 + 2 * j) : 0.0;
      const float jit = __double2float_rn(__dadd_rn((double)pot, u));
      c.m_jit[j] = jit;
      atomicMax(reinterpret_cast<int*>(c.cell_maxjit + owner), __float_as_int(jit));  // jit > 0
    }
    if (conn >= c.seg_activation_threshold) {
      atomicAdd(&c.cell_npred[owner], 1);
      if (atomicOr(&c.col_pred[owner >> 5], 1u << (owner & 31)) == 0u) atomicAdd(&c.sc[BH_SC_NPREDCOL], 1);
    }
  }
  if (b == 0 && threadIdx.x == 0) {
    c.sc[BH_SC_HAVE_PREV] = 1;
    c.sc[BH_SC_STEP] = c.sc[BH_SC_STEP] + 1;
    c.sc[BH_SC_JIT_PENDING] = want_jitter ? 0 : 1;
    if (!want_jitter) {  // no draw published M (tm_kernels.cuh, ph_activate_b)
      if (c.sc[BH_SC_X_MATCH] > c.match_capacity) atomicOr(&c.sc[BH_SC_STATUS], BH_ST_MATCH_OVERFLOW);
      c.sc[BH_SC_M] = mx;
    }
  }
}

__global__ void __launch_bounds__(BH_TM_THREADS) k_tm_shard_pack(const __grid_constant__ bh_ctx c, int* send) {
  ph_shard_pack(c, send, blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(BH_TM_THREADS) k_tm_shard_merge(const __grid_constant__ bh_ctx c, const int* recv) {
  ph_shard_merge(c, recv, blockIdx.x, gridDim.x);
}
__global__ void __launch_bounds__(BH_TM_THREADS) k_tm_activate_finish(const __grid_constant__ bh_ctx c, int want_jitter) {
  ph_activate_finish(c, blockIdx.x, gridDim.x, false, want_jitter != 0);
}
