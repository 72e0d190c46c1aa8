// placeholder (replaced by the real pipeline)
#include "../../include/alignasm_b200.h"
#include <cstring>
struct aa_ctx { char err[256]; };
extern "C" {
aa_status aa_create(aa_ctx **ctx, int) { *ctx = nullptr; return AA_ERR_NO_DEVICE; }
void aa_destroy(aa_ctx *) {}
const char *aa_last_error(const aa_ctx *) { return "stub"; }
aa_status aa_solve(aa_ctx *, const aa_batch *, const aa_opts *, aa_result *) { return AA_ERR_NO_DEVICE; }
aa_status aa_upload(aa_ctx *, const aa_batch *, aa_dev_batch **) { return AA_ERR_NO_DEVICE; }
aa_status aa_solve_device(aa_ctx *, aa_dev_batch *, const aa_opts *, aa_result *) { return AA_ERR_NO_DEVICE; }
void aa_dev_batch_free(aa_ctx *, aa_dev_batch *) {}
void aa_result_free(aa_result *) {}
const char *aa_phase_name(int) { return nullptr; }
const char *aa_version(void) { return "stub"; }
}
