/*
 * rst_b200.h -- C ABI of the B200-native stylization hot path (librst_sm100.so).
 *
 * The reference (singinwhale/realtime-style-transfer) has no FFI of its own: its boundary is the
 * Python model-builder surface.  Each entry point below names the reference
 * interface it stands in for (paths relative to the reference repository).  Plain pointers and
 * sizes only; no torch / CUDA types in the signatures (streams are passed as void* holding a
 * cudaStream_t, NULL = the legacy default stream).
 *
 * Conventions: tensors are dense NHWC float32 at the boundary, like the reference's numpy/TF
 * tensors.  "d_" pointers are device memory on the context's GPU, "h_" pointers are host memory.
 * Every call returns 0 (RST_OK) or an error code; rst_last_error() gives the message.
 * One context per GPU / per thread; no call allocates device memory after rst_commit_weights().
 */
#ifndef RST_B200_H
#define RST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rst_ctx rst_ctx;

enum rst_status {
    RST_OK = 0,
    RST_ERR_INVALID = 1,     /* bad argument / shape (the reference raises ValueError / assert) */
    RST_ERR_CUDA = 2,        /* CUDA runtime or driver failure */
    RST_ERR_STATE = 3,       /* call order violated (e.g. forward before weights committed) */
    RST_ERR_UNSUPPORTED = 4  /* valid in the reference, not built here yet */
};

enum rst_precision {
    RST_PRECISION_FP32 = 0,  /* fp32-accurate path, bar: max abs err <= 1e-4 vs the oracle.  The residual blocks' 3x3 convolutions
                              * run on tcgen05 as error-compensated split tf32 (the RST_PRECISION_TF32X3 arithmetic, fp32
                              * accumulation) when their channel counts are multiples of 32; the 9x9, strided and transposed
                              * layers and the style predictor are CUDA-core fp32 kernels */
    RST_PRECISION_BF16 = 1,  /* tcgen05 bf16 path (fp32 accumulate), bar: <= 2e-2 relative */
    RST_PRECISION_TF32 = 2,  /* tcgen05 tf32 operands on fp32 tensors (what TensorFlow itself does with fp32 convolutions on
                              * Ampere and later GPUs); operator level, the loss model and the training step only, see
                              * rst_loss_set_math / rst_train_set_math */
    RST_PRECISION_TF32X3 = 3 /* rst_op_conv2d only: error-compensated split tf32 (the arithmetic behind RST_PRECISION_FP32 in
                              * rst_loss_set_math / rst_train_set_math), for operator-level parity tests */
};

enum rst_dtype {             /* element types at the boundary of the *_typed entry points */
    RST_DTYPE_F32 = 0,
    RST_DTYPE_F16 = 1,
    RST_DTYPE_U8 = 2
};

enum rst_extractor {         /* stylePrediction.StyleFeatureExtractor, stylePrediction.py:19-22 */
    RST_EXTRACTOR_NONE = 0,
    RST_EXTRACTOR_DUMMY = 1,
    RST_EXTRACTOR_MOBILE_NET = 2
};

/* Mirrors the arguments of create_style_transfer_model (models/styleTransfer.py:213-214) plus the
 * style-predictor choice of create_style_prediction_model (models/stylePrediction.py:25-26). */
typedef struct rst_config {
    int32_t in_h, in_w, in_c;          /* content input_shape (rows, cols, G-buffer channels)     */
    int32_t out_h, out_w;              /* output_shape rows, cols (3 channels implied)            */
    int32_t bottleneck_res_y;
    int32_t bottleneck_num_filters;
    int32_t num_styles;                /* 1, or 2 for the weight-map blend (styleTransfer.py:36-44) */
    int32_t max_batch;                 /* workspace is sized for this many frames per call        */
    int32_t precision;                 /* enum rst_precision                                      */
    int32_t extractor;                 /* enum rst_extractor                                      */
    int32_t style_h, style_w;          /* style image rows, cols (3 channels implied)             */
    int32_t predictor_num_params;      /* only when in_h == 0: predictor-only context, P = this    */
} rst_config;

/* ---- lifetime -------------------------------------------------------------------------------- */
const char* rst_version(void);
/* Builds the layer plan and allocates all device workspaces (replaces the Keras graph construction in
 * styleTransfer.py:213-332 / styleTransferInferenceModel.py:9-48). */
int rst_create(const rst_config* cfg, int device, rst_ctx** out_ctx);
int rst_destroy(rst_ctx* ctx);
/* ctx may be NULL: returns the message of the last failed rst_create on this thread. */
const char* rst_last_error(const rst_ctx* ctx);

/* Page-locked host buffers for the frames / images given to the *_host entry points (what tf.data's prefetch buffers are to the
 * reference's loop, predict_video_using_checkpoint.py:90-93): asynchronous copies need pinned memory.  write_combined != 0 for
 * buffers the host only writes (G-buffers on their way to the GPU).  rst_last_error(NULL) gives the message of a failure. */
int rst_host_alloc(void** h_ptr, uint64_t num_bytes, int write_combined);
int rst_host_free(void* h_ptr);
/* Host utility: crc32c (Castagnoli) of a host buffer, the checksum of TensorFlow's tensor bundles (un-vendored third party:
 * tensorflow/core/lib/hash/crc32c; callers mask it as LevelDB does).  Used by checkpoint.py, which reads and writes the
 * reference's checkpoint format (tracing/checkpoint.py:18-37) without TensorFlow. */
uint32_t rst_host_crc32c(const void* h_data, uint64_t num_bytes);

/* ---- introspection --------------------------------------------------------------------------- */
/* num_style_parameters returned by create_style_transfer_model (styleTransfer.py:278-279, :332). */
int rst_num_style_params(const rst_ctx* ctx);
int rst_num_contract_blocks(const rst_ctx* ctx);   /* styleTransfer.py:217 */
int rst_num_expand_blocks(const rst_ctx* ctx);     /* styleTransfer.py:258 */
/* Enumerate the variables the model owns (Keras layouts: Conv2D kernel (kh,kw,in,out),
 * Conv2DTranspose kernel (kh,kw,out,in)). */
int rst_weight_count(const rst_ctx* ctx);
const char* rst_weight_name(const rst_ctx* ctx, int index);
int rst_weight_shape(const rst_ctx* ctx, int index, int64_t* shape4, int* ndim);

/* ---- weights (stands in for model.load_weights / set_weights, predict_using_checkpoint.py:84) -- */
int rst_set_weight(rst_ctx* ctx, const char* name, const float* h_data, const int64_t* shape, int ndim);
int rst_get_weight(const rst_ctx* ctx, const char* name, float* h_data, int64_t capacity_elems);
/* Uploads, folds inference BatchNorm into per-channel affines and packs the bf16 operand layouts. */
int rst_commit_weights(rst_ctx* ctx);

/* ---- the hot path ---------------------------------------------------------------------------- */
/* transfer.predict({'content','style_params'[,'style_weights']}) (predict_video_using_checkpoint.py:93-96,
 * models/styleTransfer.py:305-329).  d_content (B,in_h,in_w,in_c), d_style_params (B,S,P),
 * d_style_weights (B,out_h,out_w,S-1) or NULL when S==1, d_out (B,out_h,out_w,3) fp32 in (0,1). */
int rst_transfer_forward(rst_ctx* ctx, const float* d_content, const float* d_style_params,
                         const float* d_style_weights, float* d_out, int batch, void* stream);
/* Same with HOST buffers: H2D copies, forward, D2H copy and a stream synchronise inside. */
int rst_transfer_forward_host(rst_ctx* ctx, const float* h_content, const float* h_style_params,
                              const float* h_style_weights, float* h_out, int batch);
/* The frame loop of predict_video_using_checkpoint.py:90-98 (`for frame in dataset.prefetch(5): transfer.predict(...)`):
 * asynchronous submit of one batch with HOST buffers (pinned for real overlap).  The H2D copy of batch i+1 and the D2H
 * copy of batch i-1 overlap the forward of batch i (two staging slots).  h_out is valid after rst_transfer_wait(ticket);
 * at most two tickets are in flight (a third submit blocks until the oldest has drained). */
int rst_transfer_submit_host(rst_ctx* ctx, const float* h_content, const float* h_style_params,
                             const float* h_style_weights, float* h_out, int batch, int64_t* ticket);
int rst_transfer_wait(rst_ctx* ctx, int64_t ticket);
/* Reduced-byte variants of the three entry points above for the ends of the video loop that are not float32 in the reference
 * either: the G-buffer planes of the Unreal capture are HALF EXR channels (dataloaders/hdrScreenshots.py:14-29 widens them to
 * float32 on the host), and the reference's callers quantise the prediction at once (predict_using_checkpoint.py:99
 * `np.uint8(... * 255)`, predict_video_using_checkpoint.py:98 `(... * 255).astype(int)`).  content_dtype: RST_DTYPE_F32 or
 * RST_DTYPE_F16 (IEEE half, same NHWC layout); out_dtype: RST_DTYPE_F32 or RST_DTYPE_U8 (trunc(255*y), (B,out_h,out_w,3) bytes).
 * Style parameters and style weights stay float32.  The float32 entry points are these with F32 / F32. */
int rst_transfer_forward_typed(rst_ctx* ctx, const void* d_content, int content_dtype, const float* d_style_params,
                               const float* d_style_weights, void* d_out, int out_dtype, int batch, void* stream);
int rst_transfer_forward_host_typed(rst_ctx* ctx, const void* h_content, int content_dtype, const float* h_style_params,
                                    const float* h_style_weights, void* h_out, int out_dtype, int batch);
int rst_transfer_submit_host_typed(rst_ctx* ctx, const void* h_content, int content_dtype, const float* h_style_params,
                                   const float* h_style_weights, void* h_out, int out_dtype, int batch, int64_t* ticket);
/* style_predictor(style_image) (models/stylePrediction.py:25-75): d_style (B,style_h,style_w,3) in [0,1]
 * -> d_params (B,P). */
int rst_predict_style(rst_ctx* ctx, const float* d_style, float* d_params, int batch, void* stream);
int rst_predict_style_host(rst_ctx* ctx, const float* h_style, float* h_params, int batch);
/* inference.predict({'content','style'[,'style_weights']}) (models/styleTransferInferenceModel.py:24-39):
 * h_style (B,S,style_h,style_w,3); predictor per style, then the transfer net. */
int rst_inference_forward_host(rst_ctx* ctx, const float* h_content, const float* h_style,
                               const float* h_style_weights, float* h_out, int batch);

/* Copy one intermediate activation of the LAST forward (as fp32 NHWC) for layer-level parity tests.
 * Names follow the oracle's taps: "contract_start", "residual_block_0/conv0/relu", "expand_0", ...
 * Returns the element count through *elems; d_out may be NULL to query the size. */
int rst_debug_enable_taps(rst_ctx* ctx, int enable);
int rst_debug_tap(rst_ctx* ctx, const char* name, float* h_out, int64_t capacity_elems, int64_t* elems);
/* Number of kernels of this library launched by the last hot-path call (bench.py "gpu_launches"). */
int64_t rst_last_launch_count(const rst_ctx* ctx);
/* Average device time (ms, CUDA events on the launch stream) of named kernel groups accumulated since
 * rst_profile_reset: used by bench.py for the live roofline measurement. */
int rst_profile_enable(rst_ctx* ctx, int enable);
int rst_profile_reset(rst_ctx* ctx);
int rst_profile_get(rst_ctx* ctx, const char* group, double* total_ms, int64_t* launches);
int rst_profile_group_count(rst_ctx* ctx);
const char* rst_profile_group_name(rst_ctx* ctx, int index);

/* ---- stand-alone operators (device pointers; used by the parity tests, one reference op each) --- */
/* Message of the last failed rst_op_* call on this thread. */
const char* rst_op_last_error(void);
/* tf.keras.layers.Conv2D / Conv2DTranspose with padding='same' (styleTransfer.py:170-172, :194-199, :115-119).
 * kernel layout: Conv2D (kh,kw,ci,co); transposed!=0: (kh,kw,co,ci).  act: 0 none, 1 relu, 4 sigmoid. */
int rst_op_conv2d(const float* d_x, const float* d_kernel, const float* d_bias, float* d_y,
                  int batch, int h, int w, int ci, int co, int kh, int kw, int stride,
                  int transposed, int act, int precision, void* stream);
/* ConditionalInstanceNormalization.call (styleTransfer.py:57-71) incl. _apply_style_weights (:36-44).
 * d_params (B,S,2F) = [scale F | bias F] per style; d_weights (B,H,W,S) or NULL when S==1. */
int rst_op_cin(const float* d_x, const float* d_params, const float* d_weights, float* d_y,
               int batch, int h, int w, int f, int num_styles, int act, void* stream);
/* _apply_style_weights alone (styleTransfer.py:36-44): weights (B,H,W,2), params (B,1,2,F) -> (B,H,W,F). */
int rst_op_apply_style_weights(const float* d_weights, const float* d_params, float* d_out,
                               int batch, int h, int w, int f, void* stream);
/* get_gram_matrix_model (styleLoss.py:11-18): (B,H,W,C) -> (B,C,C), divided by H*W. */
int rst_op_gram(const float* d_x, float* d_gram, int batch, int h, int w, int c, void* stream);

/* ---- training loss (models/styleLoss.py) ------------------------------------------------------------------------ */
typedef struct rst_loss rst_loss;
/* StyleLossModelVGG(input_shape) + make_style_loss_function(loss_model, output_shape, num_styles=1, with_depth_loss=False)
 * (styleLoss.py:69-109, :295-369): VGG16 to block5_conv3 with style taps block{1,2}_conv2, block{3,4}_conv3. */
int rst_loss_create(int h, int w, int max_batch, int device, rst_loss** out);
int rst_loss_destroy(rst_loss* loss);
const char* rst_loss_last_error(const rst_loss* loss);
/* Keras VGG16 variables: "block1_conv1/kernel" (3,3,ci,co), "block1_conv1/bias", ... "block5_conv3/bias". */
int rst_loss_set_weight(rst_loss* loss, const char* name, const float* h_data, const int64_t* shape, int ndim);
int rst_loss_commit(rst_loss* loss);
/* content / style / total-variation factors (defaults 1e4, 1e-3, 1e-1: styleLoss.py:101-104) */
int rst_loss_set_factors(rst_loss* loss, float content, float style, float tv);
/* Arithmetic of the VGG16 convolutions with >= 64 input channels and of their input gradients (both on the tensor cores).
 * RST_PRECISION_FP32 (default): error-compensated split tf32, x_hi*w_hi + x_lo*w_hi + x_hi*w_lo with fp32 accumulation =
 * fp32-level accuracy.  RST_PRECISION_TF32: plain tf32 operands -- TensorFlow's own default for float32 convolutions on Ampere
 * and later -- a third of the tensor work, loss scalars still within 1e-3.  Takes effect at the next rst_loss_commit. */
int rst_loss_set_math(rst_loss* loss, int precision);
/* compute_loss(y_pred, y_true) (styleLoss.py:363-367): d_losses (B,4) = [loss, feature_loss, style_loss,
 * total_variation_loss], each a per-sample value like the reference's (B,) vectors. */
int rst_loss_forward(rst_loss* loss, const float* d_pred, const float* d_gt_content, const float* d_gt_style,
                     float* d_losses, int batch, void* stream);
/* Gradient of sum_b loss[b] w.r.t. the prediction (Keras differentiates the loss VECTOR, i.e. its batch sum:
 * styleTransferTrainingModel.py:26-29), using the activations saved by the last rst_loss_forward. */
int rst_loss_backward(rst_loss* loss, const float* d_pred, float* d_grad_pred, int batch, void* stream);

/* ---- training step (models/styleTransferTrainingModel.py + Keras Model.fit's train_step) ------------------------------------ */
typedef struct rst_trainer rst_trainer;
/* make_style_transfer_training_model (styleTransferTrainingModel.py:39-70): predictor + transfer network with num_styles = 1,
 * the VGG loss model, and tf.keras.optimizers.RMSprop state (train_network.py:102).  cfg->num_styles must be 1 and
 * cfg->extractor must name a predictor; the model always trains in fp32.  The loss model works at (out_h, out_w). */
int rst_train_create(const rst_config* cfg, int device, rst_trainer** out);
int rst_train_destroy(rst_trainer* trainer);
const char* rst_train_last_error(const rst_trainer* trainer);
/* The trainer's model / loss contexts: set variables with rst_set_weight + rst_commit_weights and rst_loss_set_weight +
 * rst_loss_commit before the first step.  The model context also serves rst_transfer_forward / rst_predict_style. */
/* Arithmetic of the step's tensor-core convolutions (the residual trunk of the transfer network, forward and input gradient,
 * and the loss model): RST_PRECISION_FP32 (default) = error-compensated split tf32 with fp32-level accuracy;
 * RST_PRECISION_TF32 = plain tf32 operands (TensorFlow's behaviour for float32 models on Ampere and later).  Everything else
 * (predictor, 9x9 / strided / transposed layers, weight gradients, normalisation) is fp32 either way.
 * Call before rst_commit_weights / rst_loss_commit. */
int rst_train_set_math(rst_trainer* trainer, int precision);
rst_ctx* rst_train_model(rst_trainer* trainer);
rst_loss* rst_train_loss(rst_trainer* trainer);
/* One forward + backward of Model.train_step: y_pred = model(x, training=True) [BatchNorm on batch statistics,
 * moving statistics updated], compute_loss (styleTransferTrainingModel.py:26-29), gradient of the batch SUM of the loss vector
 * w.r.t. every trainable variable, left in the flat gradient buffer.  d_losses (B,4) as rst_loss_forward.  d_style is the
 * predictor input (B,style_h,style_w,3); d_gt_style the style image at (out_h,out_w) for the loss model. */
int rst_train_forward_backward(rst_trainer* trainer, const float* d_content, const float* d_style, const float* d_gt_content,
                               const float* d_gt_style, float* d_losses, int batch);
/* The cudaStream_t (as void*) every kernel of the trainer is launched on: a caller that times a step with CUDA events records
 * them here (bench.py), and a data-parallel caller orders its all-reduce against it. */
void* rst_train_stream(rst_trainer* trainer);
/* Flat fp32 gradient buffer (device) holding every trainable variable's gradient: the buffer a data-parallel caller
 * all-reduces (SUM) over NCCL between rst_train_forward_backward and rst_train_apply_gradients. */
float* rst_train_gradients(rst_trainer* trainer);
int64_t rst_train_num_gradient_elements(const rst_trainer* trainer);
int rst_train_variable_range(const rst_trainer* trainer, const char* name, int64_t* offset, int64_t* elems);
/* Keras RMSprop (momentum 0, not centred): rms = rho*rms + (1-rho)*g^2; var -= lr*g/(sqrt(rms)+epsilon).
 * Defaults of RMSprop(): learning_rate 1e-3, rho 0.9, epsilon 1e-7. */
int rst_train_apply_gradients(rst_trainer* trainer, float learning_rate, float rho, float epsilon);
/* Device -> host registry and re-commit, so rst_get_weight / checkpoints / inference entry points see the trained values. */
int rst_train_sync_weights(rst_trainer* trainer);
int rst_train_read_gradient(rst_trainer* trainer, const char* name, float* h_out, int64_t capacity);
int rst_train_read_prediction(rst_trainer* trainer, float* h_out, int64_t capacity);
const float* rst_train_prediction(const rst_trainer* trainer);
/* Debug: an activation of the last step (want_grad 0) or the gradient that reached it (want_grad 1) by name:
 * "<layer>/conv", "<layer>/out", "style_params".  h_out NULL only queries the element count. */
int rst_train_debug_read(rst_trainer* trainer, const char* name, int want_grad, float* h_out, int64_t capacity, int64_t* elems);

#ifdef __cplusplus
}
#endif
#endif /* RST_B200_H */
