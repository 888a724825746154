// krig-b200: process-wide state of the C-ABI library: the last-error string, the launch counter and the pinned /
// device block pools (kb_api_internal.h says why they exist).
#include "kb_api_internal.h"

thread_local std::string g_err;
std::atomic<long long> g_launches{0};
KbPinnedPool g_pinned;
KbDevicePool g_devpool;
