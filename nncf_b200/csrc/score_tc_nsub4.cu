// score_tc_nsub4.cu — instantiates the tcgen05 score + gradient kernels for dp = 256 (all losses, both schemes).
#include "score_tc.cuh"
namespace nncf {
int launch_score_tc_nsub4(const ScoreTcArgs& a, int nblk, int R, cudaStream_t st) { return launch_score_tc_all<4>(a, nblk, R, st); }
}  // namespace nncf
