// placeholder until the tcgen05 kernel lands (next commit)
#include "common.cuh"
size_t prifit_meanshift_tc_workspace_bytes(int B, int N) { (void)B; (void)N; return 16; }
int prifit_meanshift_fwd_tc(const float*, const float*, int, int, int, float*, void*, size_t, cudaStream_t) {
    prifit_set_error("prifit_meanshift_fwd: tcgen05 engine not built");
    return PRIFIT_E_NODEVICE;
}
