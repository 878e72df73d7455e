// TEST INFRASTRUCTURE ONLY: runs the per-thread body of the second-generation InstanceNorm backward
// (ganslate_b200/csrc/instnorm_v2_core.h, the same source nvcc compiles into in_bwd_v2_kernel) on the host, one
// "thread" of one "block" at a time, with the block reduction / atomics replaced by plain sums.  Built with g++ by
// tests/test_in_bwd_v2_emul.py; nothing under ganslate_b200/ uses it.
#include "../../ganslate_b200/csrc/instnorm_v2_core.h"

namespace {

template <bool RES, int U, bool GEN>
int run(const gb_in_bwd_params& p, int cap, float neg_slope, int* grid_out) {
  bool fits = false;
  const gbv2::Geom g = gbv2::plan(p.x.N, p.x.D, p.x.H, p.x.W, p.x.C, cap, &fits);
  grid_out[0] = g.nblocks;
  grid_out[1] = g.ppb;
  const int C = p.x.C, C8 = C >> 3;
  const int slots = gbv2::slots_of(C);
  float acc1[8], acc2[8], acc3[8];
  for (int pass = 0; pass < 2; ++pass)
    for (int n = 0; n < p.x.N; ++n)
      for (int bx = 0; bx < g.nblocks; ++bx)
        for (int tid = 0; tid < gbv2::THREADS; ++tid) {
          if (pass == 0) gbv2::stream_pass<RES, U, 0, GEN>(p, g, neg_slope, tid, bx, n, acc1, acc2, acc3);
          else gbv2::stream_pass<RES, U, 1, GEN>(p, g, neg_slope, tid, bx, n, acc1, acc2, acc3);
          if (tid / C8 >= slots) continue;
          const int c = (tid % C8) * 8;
          for (int e = 0; e < 8; ++e) {
            if (pass == 0) {
              p.bstats[((int64_t)n * C + c + e) * 2 + 0] += acc1[e];
              p.bstats[((int64_t)n * C + c + e) * 2 + 1] += acc2[e];
              if (GEN && p.dprelu != nullptr) p.dprelu[c + e] += acc3[e];
            } else if (p.dbias != nullptr) {
              p.dbias[c + e] += acc1[e];
            }
          }
        }
  return 0;
}

}  // namespace

// U: pixels in flight per thread; + 10 selects the general form (PReLU, residual before the activation, scaled output)
extern "C" int in_bwd_v2_emulate(const gb_in_bwd_params* p, int cap, int U, float neg_slope, int* grid_out) {
  const bool res = p->dy_sum.ptr != nullptr;
  switch (U) {
    case 2: return res ? run<true, 2, false>(*p, cap, neg_slope, grid_out) : run<false, 2, false>(*p, cap, neg_slope, grid_out);
    case 3: return res ? run<true, 3, false>(*p, cap, neg_slope, grid_out) : run<false, 3, false>(*p, cap, neg_slope, grid_out);
    case 4: return res ? run<true, 4, false>(*p, cap, neg_slope, grid_out) : run<false, 4, false>(*p, cap, neg_slope, grid_out);
    case 12: return res ? run<true, 2, true>(*p, cap, neg_slope, grid_out) : run<false, 2, true>(*p, cap, neg_slope, grid_out);
  }
  return 1;
}

// forward: every thread of every block through gbv2::fwd_pass; U + 10 = the general form
extern "C" int in_fwd_v2_emulate(const gb_in_fwd_params* p, int cap, int U, float neg_slope) {
  bool fits = false;
  const gbv2::Geom g = gbv2::plan(p->x.N, p->x.D, p->x.H, p->x.W, p->x.C, cap, &fits);
  for (int n = 0; n < p->x.N; ++n)
    for (int bx = 0; bx < g.nblocks; ++bx)
      for (int tid = 0; tid < gbv2::THREADS; ++tid) {
        if (U == 4) gbv2::fwd_pass<4, false>(*p, g, neg_slope, tid, bx, n);
        else if (U == 2) gbv2::fwd_pass<2, false>(*p, g, neg_slope, tid, bx, n);
        else if (U == 14) gbv2::fwd_pass<4, true>(*p, g, neg_slope, tid, bx, n);
        else return 1;
      }
  return 0;
}
