// TEST INFRASTRUCTURE ONLY: host run of the on-chip InstanceNorm backward's per-thread body
// (ganslate_b200/csrc/instnorm_v3_core.h); the stash is a host array per "CTA", the warp / block reduction and the
// distributed-shared-memory exchange of a cluster are plain sums over its CTAs.  Built by tests/test_in_bwd_v2_emul.py.
#include <vector>
#include "../../ganslate_b200/csrc/instnorm_v3_core.h"

namespace {

template <bool RES, int U>
int run(const gb_in_bwd_params& p, int sms, int mode, float neg_slope, int* geom_out) {
  gbv3::Geom g;
  if (!gbv3::plan(p.x.N, p.x.D, p.x.H, p.x.W, p.x.C, sms, mode, &g)) return -1;
  geom_out[0] = g.K;
  geom_out[1] = g.ppc;
  geom_out[2] = g.steps;
  const int T = gbv3::THREADS;
  std::vector<float4> g0((size_t)g.K * g.steps * T), g1((size_t)g.K * g.steps * T);
  std::vector<uint4> xs((size_t)g.K * g.steps * T);
  for (int n = 0; n < p.x.N; ++n)
    for (int cgi = 0; cgi < p.x.C / gbv3::CG; ++cgi) {
      float tot1[gbv3::CG] = {0}, tot2[gbv3::CG] = {0};
      for (int r = 0; r < g.K; ++r)
        for (int tid = 0; tid < T; ++tid) {
          float s1[8], s2[8];
          const size_t o = (size_t)r * g.steps * T;
          gbv3::load_pass<RES, U>(p, g, neg_slope, tid, r, cgi, n, g0.data() + o, g1.data() + o, xs.data() + o, s1, s2);
          for (int e = 0; e < 8; ++e) {
            tot1[(tid % gbv3::TPP) * 8 + e] += s1[e];
            tot2[(tid % gbv3::TPP) * 8 + e] += s2[e];
          }
        }
      for (int r = 0; r < g.K; ++r)
        for (int tid = 0; tid < T; ++tid) {
          float t1[8], t2[8], db[8];
          for (int e = 0; e < 8; ++e) {
            t1[e] = tot1[(tid % gbv3::TPP) * 8 + e];
            t2[e] = tot2[(tid % gbv3::TPP) * 8 + e];
          }
          const size_t o = (size_t)r * g.steps * T;
          gbv3::apply_pass(p, g, tid, r, cgi, n, t1, t2, g0.data() + o, g1.data() + o, xs.data() + o, db);
          if (p.dbias != nullptr)
            for (int e = 0; e < 8; ++e) p.dbias[cgi * gbv3::CG + (tid % gbv3::TPP) * 8 + e] += db[e];
        }
    }
  return 0;
}

}  // namespace

// U: pixels in flight per thread; 12 / 14 = U 2 / 4 with plan mode 2 (half-size stash)
extern "C" int in_bwd_v3_emulate(const gb_in_bwd_params* p, int sms, int U, float neg_slope, int* geom_out) {
  const bool res = p->dy_sum.ptr != nullptr;
  const int mode = U >= 10 ? 2 : 1;
  if (U >= 10) U -= 10;
  if (U == 2) return res ? run<true, 2>(*p, sms, mode, neg_slope, geom_out) : run<false, 2>(*p, sms, mode, neg_slope, geom_out);
  if (U == 4) return res ? run<true, 4>(*p, sms, mode, neg_slope, geom_out) : run<false, 4>(*p, sms, mode, neg_slope, geom_out);
  return 1;
}
