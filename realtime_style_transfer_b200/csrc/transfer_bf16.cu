// bf16 tensor-core (tcgen05) transfer-network forward.  Placeholder until conv_umma.cu lands.
#include "rst_ctx.h"

namespace rst {

struct Bf16State {};
struct TrainState {};

int bf16_create(rst_ctx* ctx) { return fail(ctx, RST_ERR_UNSUPPORTED, "bf16 path not built yet"); }
int bf16_commit(rst_ctx* ctx) { return fail(ctx, RST_ERR_UNSUPPORTED, "bf16 path not built yet"); }
int bf16_transfer_forward(rst_ctx* ctx, const float*, const float*, const float*, float*, int, cudaStream_t) {
    return fail(ctx, RST_ERR_UNSUPPORTED, "bf16 path not built yet");
}
int op_conv2d_bf16(const float*, const float*, const float*, float*, int, int, int, int, int, int, int, int, int, int,
                   cudaStream_t, std::string* err) {
    *err = "bf16 path not built yet";
    return RST_ERR_UNSUPPORTED;
}

}  // namespace rst
