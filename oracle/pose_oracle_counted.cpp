// TEST INFRASTRUCTURE — the CPU restatement compiled as an exact FP64 operation counter (see counted_double.h).
// Every mpeo_* entry point exists again as mpeoc_* (same arguments: Counted is layout-compatible with double);
// mpeoc_ops_reset / mpeoc_ops_get read the counters.  Used by bench.py (op counts behind the K2 / K3 FP64 rooflines) and tests.
#include <algorithm>
#include <array>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#include "counted_double.h"
#define double Counted
#define mpeo_ mpeoc_
#include "pose_oracle_renamed.inc"
#undef double
extern "C" {
void mpeoc_ops_reset() { op_counts() = OpCounts{0, 0, 0, 0, 0, 0}; }
void mpeoc_ops_get(unsigned long long out[6]) { OpCounts& c = op_counts(); out[0] = c.add; out[1] = c.mul; out[2] = c.div; out[3] = c.sqrt; out[4] = c.cmp; out[5] = c.special; }
}
