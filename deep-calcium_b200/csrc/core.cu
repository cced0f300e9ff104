// Library plumbing: version, thread-local error text, launch counter.
#include "common.cuh"

namespace dcb {

unsigned long long g_launches = 0;

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}

// ---- dispatch policy: process-wide, read at every call (never cached), set through dcb_set_policy() ----
static const int kPolicyDefaults[DCB_POLICY_COUNT] = {
    /* FLAT */ 1, /* STRIP */ 1, /* FOLD */ 1, /* NSPLIT */ 1, /* SWAP_MIN_COUT */ 64, /* WGRAD_STRIP */ 1,
    /* BN_CTAS_PER_SM */ 4, /* PROJ_I16_SPLITS */ 0, /* SPLITK */ 1, /* FUSED_BN */ 2, /* TMA_STORE */ 1, /* BN_SLAB */ 1, /* PDL */ 0, /* PAIR */ 0};
static int g_policy[DCB_POLICY_COUNT] = {1, 1, 1, 1, 64, 1, 4, 0, 1, 2, 1, 1, 0, 0};

int policy(int key) { return (key >= 0 && key < DCB_POLICY_COUNT) ? g_policy[key] : 0; }

char* last_kernel_buf() {
  static thread_local char buf[96] = {0};
  return buf;
}
void note_kernel(const char* name) {
  char* b = last_kernel_buf();
  strncpy(b, name, 95);
  b[95] = 0;
}

}  // namespace dcb

extern "C" int dcb_set_policy(int key, int value) {
  if (key < 0 || key >= DCB_POLICY_COUNT) return dcb::fail(DCB_ERR_INVALID_ARGUMENT, "dcb_set_policy: unknown key %d", key);
  dcb::g_policy[key] = value;
  return DCB_OK;
}
extern "C" int dcb_get_policy(int key, int* value) {
  if (key < 0 || key >= DCB_POLICY_COUNT || !value) return dcb::fail(DCB_ERR_INVALID_ARGUMENT, "dcb_get_policy: bad arguments");
  *value = dcb::g_policy[key];
  return DCB_OK;
}
extern "C" int dcb_reset_policy(void) {
  for (int i = 0; i < DCB_POLICY_COUNT; ++i) dcb::g_policy[i] = dcb::kPolicyDefaults[i];
  return DCB_OK;
}
extern "C" const char* dcb_last_kernel(void) { return dcb::last_kernel_buf(); }

extern "C" int dcb_version(void) { return 200; }
extern "C" const char* dcb_last_error(void) { return dcb::last_error_buf(); }
extern "C" unsigned long long dcb_launch_count(void) { return dcb::g_launches; }
