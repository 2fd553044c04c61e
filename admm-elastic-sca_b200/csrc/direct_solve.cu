// direct_solve.cu -- placeholder while the supernodal direct solver is being brought up.
#include "common.h"
namespace admmb {
struct DirectSolver { std::vector<int> block_end; };
void direct_set_blocks(admmb_ctx *ctx, const std::vector<int> &block_end) {
	if (!ctx->direct) ctx->direct = new DirectSolver();
	ctx->direct->block_end = block_end;
}
void direct_fill_info(const admmb_ctx *ctx, admmb_info *out) { (void)ctx; (void)out; }
int direct_setup(admmb_ctx *ctx) { ADMMB_FAIL(ctx, ADMMB_E_STATE, "direct solver not built yet"); }
int direct_solve(admmb_ctx *ctx) { ADMMB_FAIL(ctx, ADMMB_E_STATE, "direct solver not built yet"); }
void direct_destroy(admmb_ctx *ctx) { delete ctx->direct; ctx->direct = nullptr; }
}
