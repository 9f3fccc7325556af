#include "engines.h"
namespace hmm {
void ring_em_run(const double *, int64_t, const HostModel &, EmResult &, cudaStream_t, hmm_info *) {
    fail(HMM_EUNSUPPORTED, "ring E/M engine not built");
}
}
