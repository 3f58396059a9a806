// TEMPORARY stubs until newsvendor.cu / netinv.cu land
#include "common.cuh"
#define STUB(name, ...) extern "C" int name(__VA_ARGS__) { orgym_set_error(#name " not implemented yet"); return ORGYM_E_UNSUPPORTED; }
STUB(orgym_newsvendor_create, const orgym_newsvendor_config_t*, int64_t, int, orgym_handle_t*)
STUB(orgym_newsvendor_destroy, orgym_handle_t)
extern "C" int64_t orgym_newsvendor_state_bytes(orgym_handle_t) { return -1; }
extern "C" int32_t orgym_newsvendor_obs_dim(orgym_handle_t) { return -1; }
STUB(orgym_newsvendor_reset, orgym_handle_t, void*, int, uint64_t, int64_t, const uint8_t*, const double*, float*, void*)
STUB(orgym_newsvendor_step, orgym_handle_t, void*, const float*, const int64_t*, int, float*, double*, uint8_t*, uint8_t*, const orgym_newsvendor_info_t*, void*)
STUB(orgym_newsvendor_export_params, orgym_handle_t, const void*, double*, void*)
STUB(orgym_newsvendor_rollout, orgym_handle_t, uint64_t, int64_t, uint32_t, const orgym_newsvendor_rollout_in_t*, const orgym_newsvendor_rollout_out_t*, void*)
STUB(orgym_netinv_create, const orgym_netinv_config_t*, int64_t, int, orgym_handle_t*)
STUB(orgym_netinv_destroy, orgym_handle_t)
extern "C" int64_t orgym_netinv_state_bytes(orgym_handle_t) { return -1; }
extern "C" int32_t orgym_netinv_obs_dim(orgym_handle_t) { return -1; }
STUB(orgym_netinv_reset, orgym_handle_t, void*, int, uint64_t, int64_t, const uint8_t*, float*, void*)
STUB(orgym_netinv_step, orgym_handle_t, void*, const float*, const double*, int, float*, double*, uint8_t*, uint8_t*, const orgym_netinv_info_t*, void*)
STUB(orgym_netinv_export_state, orgym_handle_t, const void*, double*, double*, double*, int32_t*, void*)
STUB(orgym_netinv_rollout, orgym_handle_t, void*, uint64_t, int64_t, uint32_t, const orgym_netinv_rollout_in_t*, const orgym_netinv_rollout_out_t*, void*)
