// Host-side entry points of the pipelined sweep kernel (sweep3.cu), used by the dispatcher in sweep2.cu.
#pragma once
#include "common.cuh"

namespace qio {

bool sweep3_supported(int n, int na);
size_t sweep3_workspace_bytes(int n);
size_t sweep3_mailbox_bytes(int nranks);
int sweep3_launch(double *gamma, double *S, double *U, double *W, int n, int a0, int na, const int32_t *pairs, int P,
                  const int32_t *inactive, const qio_b200_theta_tables *h_tables, qio_b200_ls_result *results,
                  double *summary, void *workspace, const qio_b200_peers *h_peers, const qio_b200_sweep_options &opts,
                  cudaStream_t st);

// the same kernel built with its phase timers (sweep3_timed.cu)
int sweep3_timed_launch(double *gamma, double *S, double *U, double *W, int n, int a0, int na, const int32_t *pairs, int P,
                        const int32_t *inactive, const qio_b200_theta_tables *h_tables, qio_b200_ls_result *results,
                        double *summary, void *workspace, const qio_b200_peers *h_peers, const qio_b200_sweep_options &opts,
                  cudaStream_t st);

}  // namespace qio
