// temporary stubs (replaced by the real ADMM / adaprox / bsdmm loops)
#include "common.cuh"
#define STUB { pmx_set_error("not implemented yet"); return PMX_ERR_UNSUPPORTED; }
extern "C" {
int pmx_nmf_adaprox_begin(pmx_nmf*, const pmx_adaprox_opts*) STUB
int pmx_nmf_adaprox_run(pmx_nmf*, int, const double*, const double*, int*, int*, int*, long long*, long long*) STUB
int pmx_nmf_bsdmm_begin(pmx_nmf*, const pmx_bsdmm_opts*) STUB
int pmx_nmf_bsdmm_run(pmx_nmf*, int, int*, int*, int*) STUB
int pmx_admm_create(pmx_ctx*, size_t, const pmx_admm_opts*, pmx_admm**) STUB
int pmx_admm_destroy(pmx_admm*) STUB
int pmx_admm_set(pmx_admm*, const float*, const float*) STUB
int pmx_admm_get(pmx_admm*, float*) STUB
int pmx_admm_init_zu(pmx_admm*) STUB
int pmx_admm_step(pmx_admm*, float, int*, int*, double*) STUB
int pmx_admm_run(pmx_admm*, float, int, int*, int*, double*) STUB
}
