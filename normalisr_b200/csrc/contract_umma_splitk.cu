// The tcgen05 contraction with work items (tile, part of the cells): contract_umma.cu compiled a second time
// with the split switched on (see the note at the top of that file).  Exports nsr_launch_contract_umma_splitk.
#define NSR_SPLITK_TU 1
#include "contract_umma.cu"
