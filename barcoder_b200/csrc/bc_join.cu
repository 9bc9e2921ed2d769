#include "bc_join.h"
bool bc_join_supported(const ComboDesc*, uint32_t) { return false; }
cudaError_t bc_join_search(JoinWorkspace&, const SearchParams&, int, cudaStream_t, cudaEvent_t, cudaEvent_t, uint32_t*) { return cudaErrorNotSupported; }
void bc_join_free(JoinWorkspace&) {}
