// score_tc_nsub2.cu — instantiates the tcgen05 score + gradient kernels for dp = 128 (all losses, both schemes).
#include "score_tc.cuh"
namespace nncf {
int launch_score_tc_nsub2(const ScoreTcArgs& a, int nblk, int R, cudaStream_t st) { return launch_score_tc_all<2>(a, nblk, R, st); }
}  // namespace nncf
