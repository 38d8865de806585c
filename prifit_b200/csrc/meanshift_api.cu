// k2 entry point: dispatch between the tcgen05 tensor-core engine and the fp32 CUDA-core engine.
#include "common.cuh"

int prifit_meanshift_fwd_simt(const float* X, const float* bw, int B, int N, int d, int T, float* newX, cudaStream_t st);
int prifit_meanshift_fwd_tc(const float* X, const float* bw, int B, int N, int T, float* newX,
                            void* ws, size_t ws_bytes, cudaStream_t st);
size_t prifit_meanshift_tc_workspace_bytes(int B, int N);

extern "C" size_t prifit_meanshift_workspace_bytes(int B, int N, int d, int engine) {
    if (engine == PRIFIT_MS_F16_TCGEN05 && d == 128) return prifit_meanshift_tc_workspace_bytes(B, N);
    return 16;
}

extern "C" int prifit_meanshift_fwd(const float* X, const float* bw, int B, int N, int d, int T, float* newX_out,
                                    int engine, void* ws, size_t ws_bytes, void* stream) {
    PF_CHECK_ARG(X && bw && newX_out, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && N > 0 && T >= 0, PRIFIT_E_BADARG, "B, N > 0 and T >= 0 required");
    if (engine == PRIFIT_MS_FP32_SIMT) return prifit_meanshift_fwd_simt(X, bw, B, N, d, T, newX_out, pf_stream(stream));
    PF_CHECK_ARG(engine == PRIFIT_MS_F16_TCGEN05, PRIFIT_E_BADARG, "unknown engine");
    PF_CHECK_ARG(d == 128, PRIFIT_E_SHAPE, "the tcgen05 engine is specialised for d == 128");
    PF_CHECK_ARG(ws && ws_bytes >= prifit_meanshift_workspace_bytes(B, N, d, engine), PRIFIT_E_WS, "workspace too small");
    return prifit_meanshift_fwd_tc(X, bw, B, N, T, newX_out, ws, ws_bytes, pf_stream(stream));
}
