// Run keys of the executor's counter-based random source (mp_rng.h); defined in mp_status.cpp.
#pragma once
#include <stdint.h>

namespace mp {
uint64_t next_run_key();   // consumes one run number (seeded) or 8 bytes of entropy (unseeded)
uint64_t peek_run_key();   // the key the next seeded run will get (mppipe_plan)
}  // namespace mp
