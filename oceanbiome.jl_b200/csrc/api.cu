// api.cu — error reporting, version and ABI introspection of libobm_b200.
#include <stdarg.h>
#include <string.h>

#include "obm_common.cuh"

namespace obm {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int launch_status(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return 0;
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return (int)e;
}

}  // namespace obm

extern "C" const char* obm_last_error(void) { return obm::g_error; }
extern "C" int obm_version(void) { return OBM_VERSION; }

extern "C" int obm_sizeof(const char* name) {
    if (!name) return OBM_ENULL;
#define S(T) \
    if (strcmp(name, #T) == 0) return (int)sizeof(T)
    S(obm_grid);
    S(obm_npd_params);
    S(obm_twoband_params);
    S(obm_multiband_params);
    S(obm_carbchem_params);
    S(obm_scale_group);
    S(obm_pisces_phyto);
    S(obm_pisces_zoo);
    S(obm_pisces_params);
    S(obm_pisces_fields);
    S(obm_sediment_params);
    S(obm_sediment_fields);
    S(obm_gas_exchange_params);
    S(obm_sugar_kelp_params);
    S(obm_particles);
    S(obm_kelp_tracers);
#undef S
    return OBM_EENUM;
}
