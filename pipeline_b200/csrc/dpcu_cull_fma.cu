// Reporting-only FMA fast mode of the cull kernel: the same source as dpcu_cull.cu compiled
// with -fmad=true, so ptxas may contract a*b+c into one rounding.  NOT bit-exact with the
// reference; bench.py / tests report the boundary objects on which it disagrees
// (BASELINE.json north_star: "any fused-multiply-add fast mode must report its
// boundary-object disagreements separately").  Selected with DPCU_CULL_OPT_FMA = 1.
#define DPCU_FMA_VARIANT 1
#include "dpcu_cull.cu"
