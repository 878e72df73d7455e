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
