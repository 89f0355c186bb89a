// K3/K4: supernodal sparse Cholesky (placeholder until the numeric phase lands).
#include "nes_internal.h"

namespace nes {

int sparse_analyze(nes_ctx* c, nes_matrix*, nes_factor*) {
    return fail(c, NES_ERR_INVALID, "sparse Cholesky is not implemented yet");
}
int sparse_factorize(nes_ctx* c, nes_matrix*, nes_factor*) {
    return fail(c, NES_ERR_INVALID, "sparse Cholesky is not implemented yet");
}
int sparse_solve_inplace(nes_ctx* c, nes_factor*, double*) {
    return fail(c, NES_ERR_INVALID, "sparse Cholesky is not implemented yet");
}
void sparse_free(nes_ctx*, nes_factor*) {}

}  // namespace nes
