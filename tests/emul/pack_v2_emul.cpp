// TEST INFRASTRUCTURE ONLY: host run of the second-generation pack / unpack element bodies
// (ganslate_b200/csrc/pack_v2_core.h); built by tests/test_pack_v2_emul.py.
#include "../../ganslate_b200/csrc/pack_v2_core.h"

extern "C" int pack_v2_emulate(const gb_pack_params* p) {
  for (int cls = 0; cls < p->nclass; ++cls) {
    const uint32_t total8 = (uint32_t)p->rows_pad * ((uint32_t)p->kpad[cls] >> 3);
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p->dst) + p->w_offset[cls]);
    for (uint32_t i8 = 0; i8 < total8; ++i8) dst[i8] = gbp2::pack8(*p, cls, i8);
  }
  return 0;
}

extern "C" int unpack_v2_emulate(const gb_unpack_batch* b) {
  for (int k = 0; k < b->count; ++k) {
    const gb_unpack_item& it = b->item[k];
    const uint32_t total = (uint32_t)it.rows * (uint32_t)it.chans * (uint32_t)it.ntaps;
    for (uint32_t i = 0; i < total; ++i) gbp2::unpack1(it, i);
  }
  return 0;
}
