// quisk_b200/csrc/rxfused.cu -- fused shared-memory cascade for the full-rate decimator (placeholder
// until the cascade kernel lands: the chain then runs stage by stage).
#include "rxchain.h"

namespace qc {

bool RxChain::fused_applicable() { return false; }
int RxChain::run_fused_decimator(const cd *, long, int, cd *, long, int *, cudaStream_t) { set_error("fused decimator not built"); return QC_EINVAL; }
int RxChain::reset_fused() { return QC_OK; }
void RxChain::release_fused() {}

}  // namespace qc
