// Link-time stand-ins for the CUDA half of the library: the sanitizer build exercises host code only.
#include "binder_b200.h"
#include <cstdlib>
extern "C" { void* bb_host_alloc(size_t n) { return malloc(n ? n : 1); } void bb_host_free(void* p) { free(p); }
uint32_t bb_engine_max_batch(const bb_engine*) { return 1u << 20; } uint32_t bb_engine_max_batch_bytes(const bb_engine*) { return 1u << 26; }
int bb_resolve_batch(bb_engine*, const uint8_t*, const uint32_t*, uint32_t, uint64_t, uint32_t, uint8_t*, uint32_t, uint32_t*, uint16_t*, uint8_t*, uint32_t*, uint32_t*) { return BB_ERR_NO_DEVICE; } }
