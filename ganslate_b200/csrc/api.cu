// Error plumbing, version and bring-up knobs of the C ABI.
#include "gb_common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";
int g_gb_knobs[32] = {0};
unsigned long long g_gb_launches = 0;

void gb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* gb_last_error(void) { return g_err; }
extern "C" int gb_version(void) { return GB_VERSION; }
extern "C" int gb_debug_knob(int knob, int value) {
  if (knob < 0 || knob >= 32) return -1;
  const int old = g_gb_knobs[knob];
  g_gb_knobs[knob] = value;
  return old;
}
extern "C" unsigned long long gb_launch_count(void) { return g_gb_launches; }

extern "C" int gb_workspace_bytes(int op, const void* params, int64_t* bytes) {
  GB_CHECK(params != nullptr && bytes != nullptr, "gb_workspace_bytes: null pointer");
  if (op == GB_WS_WGRAD) {
    const gb_wgrad_params* p = static_cast<const gb_wgrad_params*>(params);
    GB_CHECK(p->rows >= 1 && p->kpad >= 64 && p->kpad % 64 == 0, "gb_workspace_bytes: bad rows / kpad %d / %d", p->rows, p->kpad);
    *bytes = (int64_t)((p->rows + 127) / 128 * 128) * p->kpad * 4;
    return 0;
  }
  if (op == GB_WS_IN_BWD) {
    const gb_view* x = static_cast<const gb_view*>(params);
    *bytes = ((int64_t)x->N * x->C * 2 + 4) * 4;
    return 0;
  }
  GB_CHECK(false, "gb_workspace_bytes: unknown operator %d", op);
}
