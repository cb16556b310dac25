// libmarsgb: instantiation unit of the dense pre-aggregation kernels for one (NK, THREADS) pair;
// compiled four times by mars_b200/_build.py with -DDN_NK={1,2} -DDN_TH={256,512}.
#include "mgb_dense.cuh"

#define CAT_(a, b, c, d) a##b##c##d
#define CAT(a, b, c, d) CAT_(a, b, c, d)
int CAT(mgb_dense_launch_, DN_NK, _, DN_TH)(const DenseParams& p, const DenseBatch& bt, int ncol, bool plain, int grid,
                                              size_t smem, cudaStream_t st) {
  return mgb_dense_launch<DN_NK, DN_TH>(p, bt, ncol, plain, grid, smem, st);
}
