// a16: Bloch-phase periodic elimination (src/assemble_maxwell.cpp:496-574), sparse.
#include "common.cuh"
using namespace efb;
extern "C" {
int efb_periodic_extra(const efb_mesh *, int64_t, const int32_t *, const int32_t *, int32_t, const int32_t *, const int32_t *, int64_t *, int32_t *, int32_t *) {
  return fail(nullptr, EFB_ERR_STATE, "efb_periodic_extra: not implemented yet");
}
int efb_apply_periodic(efb_system *sys_, int32_t, int32_t, int32_t, const int32_t *, const int32_t *, const double *) {
  return fail(sys_ ? ((System *)sys_)->ctx : nullptr, EFB_ERR_STATE, "efb_apply_periodic: not implemented yet");
}
}
