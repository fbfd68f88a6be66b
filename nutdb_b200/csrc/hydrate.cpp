// placeholder, replaced below
#include "../../include/nutdb_gpu.h"
extern "C" size_t nutdb_fmt_debug(const NutdbBatch*, uint64_t, const uint8_t*, size_t, char*, size_t) { return 0; }
extern "C" size_t nutdb_fmt_error(const NutdbBatch*, uint64_t, const uint8_t*, size_t, char*, size_t) { return 0; }
