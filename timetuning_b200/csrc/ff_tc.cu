// FF stage 2, TC engine (tcgen05 / TMEM / TMA) — placeholder until the kernel lands.
#include "common.cuh"

namespace timet {

bool ff_tc_supported(const timet_ff_params &p) { (void)p; return false; }

int ff_select_tc_launch(const timet_ff_params &p, const FFLayout &L, char *ws, cudaStream_t st) {
    (void)p; (void)L; (void)ws; (void)st;
    set_error("tensor-core engine not built");
    return TIMET_ERR_UNSUPPORTED;
}

}  // namespace timet
