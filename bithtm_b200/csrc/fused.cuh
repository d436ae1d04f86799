// The whole SP+TM timestep (HierarchicalTemporalMemory.process, networks.py:146-149)
// as ONE kernel: the phase functions of sp_kernels.cuh / tm_kernels.cuh separated by
// barriers.  MODE 1: the grid is a single thread-block cluster (hardware
// barrier.cluster, for networks whose step is latency-bound); MODE 2: a cooperative
// grid with one CTA per SM and a global-memory barrier (HBM-bound sizes).
// One launch can run several consecutive steps from the device input ring; such launches of large networks use
// k_step_pipe (end of this file): the same phases as two pipelines on two teams of CTAs.
#pragma once

#include "sp_kernels.cuh"
#include "tm_kernels.cuh"

#define FUSED_THREADS 1024

template <int MODE>
__global__ void __launch_bounds__(FUSED_THREADS, 1)
    k_step_fused(const __grid_constant__ bh_ctx c, const uint32_t* input_fixed, int n_steps, int flags, int want_summary) {
  extern __shared__ __align__(16) uint32_t s_dyn[];
  // flags: BH_STEP_LEARNING | BH_STEP_NO_WINNER_CELLS (TemporalMemory.process's learning / return_winner_cell,
  // networks.py:91): winner cells are formed (and rand(k, c) drawn) when learning or asked for (:99); the
  // jitter of the activation is drawn only with return_winner_cell (:121) -- else it stays pending until a
  // later step needs it.  Neither: the inference-only step, which draws nothing and updates only duty cycles.
  const int learning = flags & BH_STEP_LEARNING;
  const bool want_jit = !(flags & BH_STEP_NO_WINNER_CELLS);
  const bool want = learning || want_jit;
  const int b = blockIdx.x, nb = gridDim.x;
  const int nw = nb > 1 ? nb - 1 : 1;  // CTAs running the ranged TM phases
  const bool worker = b < nw;
  const bool rng = b == nb - 1;        // CTA producing the random draws
  // Small networks (cluster mode): independent pieces of a phase run on DIFFERENT CTAs instead of one after
  // the other on all of them -- CTAs [0, nsel) form the winner lists while the others learn the SP rows
  // (P2) and flag the learning segments (P3), so a phase costs its longest chain of dependent L2 round
  // trips, not their sum.  Partitions: select (b, nsel); learn-select (b - nl0, nlrn); scan (b, nw).
  const bool split = MODE == 1 && nw >= 8;
  const int nsel = split ? 4 : nw;
  const int nl0 = split ? nsel : 0;
  const int nlrn = nw - nl0;
  GridBar bar = grid_bar_open(reinterpret_cast<unsigned int*>(c.sc + BH_SC_BAR_COUNT));
#define BH_SYNC()                          \
  do {                                     \
    if (MODE == 1) cluster_barrier();      \
    else grid_barrier(bar, (unsigned)nb);  \
  } while (0)

  // phase timestamps of the last step (CTA 0): ctx.blk row 7, as 64-bit globaltimer ns
  unsigned long long* stamps = reinterpret_cast<unsigned long long*>(c.blk + 7 * BH_BLK_STRIDE);
  int stamp_i = 0;
#define BH_STAMP()                                                            \
  do {                                                                        \
    if (b == 0 && threadIdx.x == 0) {                                         \
      unsigned long long t_;                                                  \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                  \
      stamps[stamp_i] = t_;                                                   \
    }                                                                         \
    ++stamp_i;                                                                \
  } while (0)
  const int pos0 = c.sc[BH_SC_INPUT_POS];
  // want_summary == 2: zero-copy host step -- `input_fixed` is pinned HOST memory (read over PCIe once
  // per CTA in P0 and staged into input_dev for the learning phase) and summary_dev is pinned HOST
  // memory too, completed by a flag word the host spins on
  const bool zero_copy = want_summary == 2;
  for (int step = 0; step < n_steps; ++step) {
    const uint32_t* input =
        input_fixed ? input_fixed : c.input_ring + (long long)((pos0 + step) % c.ring_len) * c.input_words;
    stamp_i = 0;
    BH_STAMP();
    // P0: overlap + boost on all CTAs; draw #1 (rand(k, c)) on the rng CTA
    if (rng && want) {
      ph_fill_jitter(c, s_dyn);  // the previous activation's deferred rand(M) first (no-op otherwise)
      __syncthreads();
      ph_draw(c, 1, 1, nw);
    }
    constexpr bool HIST = MODE == 2;  // the grid-wide selection starts from a histogram built here
    if (nb == 1) ph_overlap<true, HIST>(c, input, s_dyn, 0, 1);
    else if (!rng) ph_overlap<true, HIST>(c, input, s_dyn, b, nb - 1);  // the rng CTA is busy drawing
    if (zero_copy) {
      if (b == 0)  // s_dyn holds the input words (ph_overlap staged them)
        for (int i = threadIdx.x; i < c.input_words; i += blockDim.x) c.input_dev[i] = s_dyn[i];
      input = c.input_dev;
    }
    BH_SYNC();
    BH_STAMP();
    // P1: global inhibition (one CTA)
    if (MODE == 2 && c.column_dim >= 16384) {  // grid-wide selection (its own barriers inside)
      if (b == 0) retire_prev_flags(c);
      topk_grid_hist(c, reinterpret_cast<const unsigned long long*>(c.boosted), c.column_dim, c.active_columns,
                     c.active_cols + (c.sc[BH_SC_STEP] & 1) * c.active_columns, c.col_active, b, nb, bar);
    } else if (b == 0) {
      ph_topk(c);
    }
    if (rng && nb > 1) ph_rng_speculate(c, 2);  // idle here: produce half of the stream words this step will draw
    BH_SYNC();
    BH_STAMP();
    // P2: SP learning + duty cycles; bursting / winner bits per active column.
    // (Measured and rejected at cfg3: the P2..P4 bookkeeping chain on a TEAM of CTAs next to the SP learning pass
    // -- 96 us against 64 + 23, its dependent loads queue behind the bandwidth-bound traffic; and the P3 + P4
    // chain on ONE CTA in a single phase -- 40 us against 23.)
    if (split) {
      if (b < nsel) {
        ph_select_a(c, b, nsel, want);
      } else {
        if (learning) ph_sp_learn<true>(c, input, b - nsel, nb - nsel);
        ph_duty(c, b - nsel, nb - nsel);
      }
    } else {
      if (MODE == 2 && rng && c.column_dim >= 16384) tk3_rebin_from_selection(c);  // (no-op unless the selection fell back)
      if (learning) ph_sp_learn<MODE == 1>(c, input, b, nb);
      ph_duty(c, b, nb);
      if (worker) ph_select_a(c, b, nw, want);
    }
    BH_SYNC();
    BH_STAMP();
    // P3: ordered winner lists; learning / punished flags among previous matching segments
    if (b < nsel) ph_select_b(c, b, nsel, want);
    if (b >= nl0 && worker) ph_learn_select_a(c, learning, b - nl0, nlrn);
    if (rng && nb > 1) ph_rng_speculate(c, 1);  // idle here too: the other half
    BH_SYNC();
    BH_STAMP();
    // P4: learning lists, recycled / new segments; draw #2 (rand(L, W+1)) on the rng CTA
    if (rng) ph_draw(c, 2, learning, nlrn, MODE == 2);
    if (b >= nl0 && worker) ph_learn_select_b(c, learning, b - nl0, nlrn);
    BH_SYNC();
    BH_STAMP();
    // A LAZY step (mt19937.cuh; decided by draw #2, published by the barrier) does not materialise
    // rand(L, W+1): stage 1 of the learning pass finds the rows that grow, jumps produce their words
    // and the words after the matrix.
    const bool lazy = MODE == 2 && c.rng64[R_LAZY] != 0;
    // P4b: the stream words draw #2 planned for many CTAs (mt19937.cuh)
    if (MODE == 2 && c.jump_polys > 0 && !lazy) {
      ph_rng_chunks(c, s_dyn, b, nb);
      BH_SYNC();
    }
    BH_STAMP();
    // P5: permanence updates, deletion, growth
    if (MODE == 2 && lazy) {
      // stage 1 of the learning pass and the tail jumps, both spread over every CTA (measured and rejected:
      // 48 CTAs of stage 1 next to the jumps on the rest -- no gain on one GPU, and the jump units then need more
      // rounds on a shard: 28 us against 16 at 8 shards)
      if (learning) ph_learn_apply(c, s_dyn, b, nb, 1);
      ph_rng_jumps(c, s_dyn, 0, (int)c.rng64[R_TAIL_CHUNKS], 0, 0, b, nb);
      BH_SYNC();
      ph_rng_lazy_rows(c, s_dyn, b, nb, [&]() { BH_SYNC(); },
                       [&](bool produce_rows) { ph_learn_grow(c, s_dyn, b, nb, produce_rows); });
    } else {
      if (learning) ph_learn_apply(c, s_dyn, b, nb);
      BH_SYNC();
    }
    BH_STAMP();
    // (P6, once a phase of its own -- retiring the previous activation words, the winner index, the segment
    // count -- now runs at the head of P7: the activation words are double-buffered, so nothing it writes is
    // read by the scan)
    BH_STAMP();
    // P7: segment potentials; the drawing CTA is idle for the whole scan: it produces the stream words of
    // the rest of this step and of the next one (the P1 / P3 calls then only top up)
    // (a lazy step: its last CTAs generate the words after the skipped matrix instead of scanning)
    const int ns = (MODE == 2 && lazy) ? nb - rng_tail_ctas(c, nb) : nw;  // CTAs that scan the segments
    const bool scanner = b < ns;
    if (MODE == 2 && lazy) ph_rng_lazy_tail(c, s_dyn, b, nb);
    ph_post(c, b, nb);
    if (scanner) ph_activate_a(c, b, ns);
    if (rng && nb > 1 && !lazy) ph_rng_speculate(c, 1, true);
    BH_SYNC();
    BH_STAMP();
    // P8: draw #3 (rand(M)) -- a phase of its own only when the words draw #2 left produced do not
    // cover it (every CTA takes the same decision from barrier-published values)
    bool ready3;
    int m_before, m_total;  // this CTA's offset in the matching list and its length (reused by P9)
    {
      __shared__ int s_red3[32];
      blk_prefix(BLK(c, BLK_MATCH), scanner ? b : 0, ns, s_red3, m_before, m_total);
      ready3 = !want_jit || (long long)(m_total < c.match_capacity ? m_total : c.match_capacity) <= c.rng64[R_READY3];
      __syncthreads();
    }
    if (!ready3) {
      if (rng) ph_draw(c, 3, 1, ns);
      BH_SYNC();
    }
    BH_STAMP();
    // P9: matching list, jitter, predictions; completes the step
    if (want_jit && ready3 && rng) ph_draw3_ready(c, ns, m_total);
    if (scanner) ph_activate_b(c, b, ns, ready3, want_jit, m_before, m_total);
    BH_SYNC();
    BH_STAMP();
  }
  if (want_summary) ph_summary(c, b, nb);
  if (zero_copy) {  // every CTA's summary words are on their way to host memory; then the flag
    __threadfence_system();
    BH_SYNC();
    if (b == 0 && threadIdx.x == 0) {
      volatile int* flag = c.summary_dev + BH_SUMMARY_INTS(c.active_columns);
      *flag = c.sc[BH_SC_STEP];  // completed steps, >= 1
      __threadfence_system();
    }
  }
  if (!input_fixed && b == 0 && threadIdx.x == 0) c.sc[BH_SC_INPUT_POS] = pos0 + n_steps;
#undef BH_SYNC
#undef BH_STAMP
}

// ---------------------------------------------------------------------------------------------------------
// The same step as a TWO-PIPELINE cooperative kernel (large networks, several steps per launch).
// The spatial pooler of step s+1 does not depend on the temporal memory of step s: its inputs are the SP state
// after step s (permanence, mask, duty cycles) and the next input.  So the grid is split into two teams with
// their own barriers that meet once per step:
//   SP team (first CTAs):  learn(s) + duty(s) | overlap(s+1) + histogram | selection(s+1) -> a staging list
//   TM team (last CTAs) :  draw 1 | winner bits | lists | learning lists + draw 2 | learn | scan | jitter   (step s)
//   join; ONE CTA commits the staged active columns of step s+1 (flags, list); join.
// The HBM-bound SP passes and the latency-bound TM chain then overlap instead of adding up.  The first step
// of a launch computes its overlap + selection on the whole grid, the last one runs no SP front, so what a
// launch leaves behind (State fields, stream position, learned state) is exactly what k_step_fused<2> leaves.
// ctx.pipe_ctas = CTAs of the TM team.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FUSED_THREADS, 1)
    k_step_pipe(const __grid_constant__ bh_ctx c, const uint32_t* input_fixed, int n_steps, int flags, int want_summary) {
  extern __shared__ __align__(16) uint32_t s_dyn[];
  const int learning = flags & BH_STEP_LEARNING;
  const bool want_jit = !(flags & BH_STEP_NO_WINNER_CELLS);
  const bool want = learning || want_jit;
  const int b = blockIdx.x, nb = gridDim.x;
  // Team sizes are chosen per step (published by the committing CTA, below): ctx.pipe_ctas CTAs for the TM team
  // in steady state -- its phases are dependent round trips, a few dozen SMs carry them -- and 70 % of the grid
  // after a step in which many learning segments grew (growth is one CTA per row: the start of a fresh network).
  // (Measured and rejected: such steps with every CTA playing both roles one after the other -- the role
  // variables cost registers, ptxas spilled inside the scan's loop, 60 -> 80 us; and the phases out of line.)
  const int nt_light = c.pipe_ctas, nt_heavy = nt_light > nb - nb * 3 / 10 ? nt_light : nb - nb * 3 / 10;
  int nt = nt_light, ns = nb - nt;
  bool sp_team = b < ns;
  int tb = b - ns;                      // index inside the TM team
  const bool rng = b == nb - 1;         // CTA producing the random draws (always in the TM team)
  GridBar barA = grid_bar_open(reinterpret_cast<unsigned int*>(c.sc + BH_SC_BAR_COUNT));
  GridBar barT = grid_bar_open(reinterpret_cast<unsigned int*>(c.sc + BH_SC_BAR2_COUNT));
  GridBar barS = grid_bar_open(reinterpret_cast<unsigned int*>(c.sc + BH_SC_BAR3_COUNT));
  unsigned long long* stamps = reinterpret_cast<unsigned long long*>(c.blk + 7 * BH_BLK_STRIDE);
  // stamps of the last step: SP team's CTA 0 at [0..], TM team's first CTA at [64..]
#define BH_STAMP_AT(i)                                             \
  do {                                                             \
    if (threadIdx.x == 0 && (b == 0 || tb == 0) && stamp_it) {     \
      unsigned long long t_;                                       \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));       \
      stamps[(b == 0 ? 0 : 64) + (i)] = t_;                        \
    }                                                              \
  } while (0)
  const int step0 = c.sc[BH_SC_STEP];
  const int pos0 = c.sc[BH_SC_INPUT_POS];
  const int k = c.active_columns;
  const bool zero_copy = want_summary == 2;
  bool stamp_it = true;  // (the stamps describe the last PIPELINED iteration of a launch)
  bool drew1 = false;    // draw #1 of the coming step was taken at the end of the previous one
  int* stage = c.active_cols + 2 * k;  // the selection of the step ahead
  const unsigned long long* keys = reinterpret_cast<const unsigned long long*>(c.boosted);

  // ---- front of the first step on the whole grid: overlap + histogram, selection
  {
    const uint32_t* input = input_fixed ? input_fixed : c.input_ring + (long long)(pos0 % c.ring_len) * c.input_words;
    BH_STAMP_AT(0);
    ph_overlap<true, true>(c, input, s_dyn, b, nb, step0);
    if (zero_copy && b == 0)  // s_dyn holds the input words (ph_overlap staged them)
      for (int i = threadIdx.x; i < c.input_words; i += blockDim.x) c.input_dev[i] = s_dyn[i];
    grid_barrier(barA, nb);
    BH_STAMP_AT(1);
    if (b == 0) {
      retire_prev_flags(c);
      if (threadIdx.x == 0)  // (the TM team is idle here: the counters of the previous step are final)
        c.sc[BH_SC_PIPE_SPLIT] = (c.sc[BH_SC_L] > 0 && c.sc[BH_SC_NGROW] * 8 > c.sc[BH_SC_L]) ? nt_heavy : nt_light;
    }
    topk_grid_hist(c, keys, c.column_dim, k, c.active_cols + (step0 & 1) * k, c.col_active, b, nb, barA, step0);
  }
  for (int step = 0; step < n_steps; ++step) {
    const int s = step0 + step;
    const bool more = step + 1 < n_steps;
    stamp_it = more || n_steps == 1;
    const uint32_t* input =
        zero_copy ? c.input_dev
                  : (input_fixed ? input_fixed : c.input_ring + (long long)((pos0 + step) % c.ring_len) * c.input_words);
    grid_barrier(barA, nb);  // the active columns of step s are committed; step s-1 is complete
    if (c.sc[BH_SC_PIPE_SPLIT] != nt) {  // (stable until the next commit; identical on every CTA)
      nt = c.sc[BH_SC_PIPE_SPLIT];
      ns = nb - nt;
      sp_team = b < ns;
      tb = b - ns;
      barT = grid_bar_open(reinterpret_cast<unsigned int*>(c.sc + BH_SC_BAR2_COUNT));  // (no team barrier is in flight)
      barS = grid_bar_open(reinterpret_cast<unsigned int*>(c.sc + BH_SC_BAR3_COUNT));
    }
    BH_STAMP_AT(2);
    if (sp_team) {
      if (b == ns - 1) tk3_rebin_from_selection(c, nullptr, s);  // (no-op unless the selection fell back)
      if (learning) ph_sp_learn<false>(c, input, b, ns, s);
      ph_duty(c, b, ns);
      BH_STAMP_AT(3);
      if (more) {
        const uint32_t* next = input_fixed ? input_fixed : c.input_ring + (long long)((pos0 + step + 1) % c.ring_len) * c.input_words;
        grid_barrier(barS, ns);
        ph_overlap<true, true>(c, next, s_dyn, b, ns, s + 1);
        grid_barrier(barS, ns);
        BH_STAMP_AT(4);
        topk_grid_hist(c, keys, c.column_dim, k, stage, nullptr, b, ns, barS, s + 1);
        BH_STAMP_AT(5);
      }
    } else {
      const int nw = nt - 1;                 // TM CTAs running the ranged phases
      const bool worker = tb < nw;
      // ---- temporal memory of step s (phases as in k_step_fused<2>, on nt CTAs)
      BH_STAMP_AT(0);
      if (!drew1) {  // (else the drawing CTA took draw #1 at the end of the previous step, see below)
        if (rng && want) {
          ph_fill_jitter(c, s_dyn);
          __syncthreads();
          ph_draw(c, 1, 1, nw);
        }
        grid_barrier(barT, nt);
      }
      BH_STAMP_AT(1);
      if (worker) ph_select_a(c, tb, nw, want);
      if (rng) ph_rng_speculate(c, 2);
      grid_barrier(barT, nt);
      BH_STAMP_AT(2);
      if (worker) {
        ph_select_b(c, tb, nw, want);
        ph_learn_select_a(c, learning, tb, nw);
      }
      if (rng) ph_rng_speculate(c, 1);
      grid_barrier(barT, nt);
      BH_STAMP_AT(3);
      if (rng) ph_draw(c, 2, learning, nw, true);
      if (worker) ph_learn_select_b(c, learning, tb, nw);
      grid_barrier(barT, nt);
      BH_STAMP_AT(4);
      const bool lazy = c.rng64[R_LAZY] != 0;
      if (c.jump_polys > 0 && !lazy) {
        ph_rng_chunks(c, s_dyn, tb, nt);
        grid_barrier(barT, nt);
      }
      BH_STAMP_AT(5);
      if (lazy) {
        if (learning) ph_learn_apply(c, s_dyn, tb, nt, 1);
        ph_rng_jumps(c, s_dyn, 0, (int)c.rng64[R_TAIL_CHUNKS], 0, 0, tb, nt);
        grid_barrier(barT, nt);
        ph_rng_lazy_rows(c, s_dyn, tb, nt, [&]() { grid_barrier(barT, nt); },
                         [&](bool produce_rows) { ph_learn_grow(c, s_dyn, tb, nt, produce_rows); });
      } else {
        if (learning) ph_learn_apply(c, s_dyn, tb, nt);
        grid_barrier(barT, nt);
      }
      BH_STAMP_AT(6);
      const int nscan = lazy ? nt - rng_tail_ctas(c, nt) : nw;  // CTAs that scan the segments
      const bool scanner = tb < nscan;
      if (lazy) ph_rng_lazy_tail(c, s_dyn, tb, nt);
      ph_post(c, tb, nt);
      if (scanner) ph_activate_a(c, tb, nscan);
      if (rng && !lazy) ph_rng_speculate(c, 1, true);
      grid_barrier(barT, nt);
      BH_STAMP_AT(7);
      bool ready3;
      int m_before, m_total;
      {
        __shared__ int s_red3[32];
        blk_prefix(BLK(c, BLK_MATCH), scanner ? tb : 0, nscan, s_red3, m_before, m_total);
        ready3 = !want_jit || (long long)(m_total < c.match_capacity ? m_total : c.match_capacity) <= c.rng64[R_READY3];
        __syncthreads();
      }
      if (!ready3) {
        if (rng) ph_draw(c, 3, 1, nscan);
        grid_barrier(barT, nt);
      }
      BH_STAMP_AT(8);
      if (want_jit && ready3 && rng) ph_draw3_ready(c, nscan, m_total);
      if (scanner) ph_activate_b(c, tb, nscan, ready3, want_jit, m_before, m_total);
      // the next step's rand(k, c) follows this step's rand(M) in the stream: with the jitter drawn every step
      // (nothing can be pending) the drawing CTA takes it now, and the next step starts with its winner bits
      if (more && want_jit && rng) ph_draw(c, 1, 1, nw);
      BH_STAMP_AT(9);
    }
    drew1 = more && want_jit;
    if (more) {
      grid_barrier(barA, nb);
      BH_STAMP_AT(sp_team ? 6 : 10);
      if (b == 0) {  // commit the staged selection: flags of step s retired, those of step s+1 set, the list copied
        if (threadIdx.x == 0)
          c.sc[BH_SC_PIPE_SPLIT] = (c.sc[BH_SC_L] > 0 && c.sc[BH_SC_NGROW] * 8 > c.sc[BH_SC_L]) ? nt_heavy : nt_light;
        retire_prev_flags(c);  // (the step counter already says s+1: "previous" is the list of step s)
        int* out = c.active_cols + ((s + 1) & 1) * k;
#pragma unroll 1
        for (int i = threadIdx.x; i < k; i += blockDim.x) {
          const int col = stage[i];
          out[i] = col;
          c.col_active[col] = 1;
        }
      }
    }
  }
  grid_barrier(barA, nb);
  if (want_summary) ph_summary(c, b, nb);
  if (zero_copy) {
    __threadfence_system();
    grid_barrier(barA, nb);
    if (b == 0 && threadIdx.x == 0) {
      volatile int* flag = c.summary_dev + BH_SUMMARY_INTS(c.active_columns);
      *flag = c.sc[BH_SC_STEP];
      __threadfence_system();
    }
  }
  if (!input_fixed && b == 0 && threadIdx.x == 0) c.sc[BH_SC_INPUT_POS] = pos0 + n_steps;
#undef BH_STAMP_AT
}
