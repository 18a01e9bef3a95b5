// XLA typed-FFI handlers over the plain C ABI of include/a2m.h ("XLA FFI flavour"): the custom calls a JAX host
// binds with jax.ffi.register_ffi_target / jax.ffi.ffi_call in place of
//     jax.vmap(model.predict, in_axes=(None, 0, None))(state, samples, rope_freqs)                 (reference infer.py:40)
//     eqx.filter_value_and_grad(compute_loss)(model, state, audio, rope_freqs, events, scale, key)  (reference train.py:48-62)
//     tx.update + eqx.apply_updates                                                                 (reference train.py:324-325)
//
// NOT BUILT IN THIS IMAGE: xla/ffi/api/ffi.h ships with jaxlib, which is absent, so this file has never been
// compiled or run.  audio-to-midi_b200/build.py:build_xla_ffi() compiles it (g++, no nvcc needed: it only forwards pointers)
// into _build/liba2m_xla_ffi.so when `import jax.ffi` resolves; audio-to-midi_b200/jax_binding.py registers the targets.
//
// Contract kept by every handler (XLA may call one per device from different host threads):
//   * everything is enqueued on the stream XLA passes (ffi::PlatformStream<cudaStream_t>); no host synchronisation;
//   * every buffer is XLA-owned device memory, never retained or freed here; scratch lives in the A2mHandle;
//   * errors come back as ffi::Error with a2m_last_error's text; nothing throws or aborts across the boundary;
//   * the only global state is the read-only `handle` attribute: the A2mHandle* (as an int64) the Python side created for the
//     device the computation is placed on -- one handle per device, a handle is used by one computation at a time.
#include <cstdint>

#include <cuda_runtime_api.h>

#include "xla/ffi/api/c_api.h"
#include "xla/ffi/api/ffi.h"

#include "../../include/a2m.h"

namespace ffi = xla::ffi;

namespace {

inline A2mHandle* as_handle(int64_t v) { return reinterpret_cast<A2mHandle*>(static_cast<intptr_t>(v)); }

inline ffi::Error status(A2mHandle* h, int rc, const char* what) {
  if (rc == A2M_OK) return ffi::Error::Success();
  return ffi::Error(rc == A2M_EINVAL ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                    std::string(what) + ": " + a2m_last_error(h));
}

inline bool is_windows(const ffi::Buffer<ffi::F32>& audio) {
  auto d = audio.dimensions();
  return d.size() == 3 && d[1] == 2 && d[2] == A2M_WINDOW_SAMPLES;
}

// ---- a2m_forward: (audio [B,2,80000], cos [P,32], sin [P,32]) -> (logits [B,250,90], probs [B,250,90])
ffi::Error ForwardImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> audio, ffi::Buffer<ffi::F32> cos, ffi::Buffer<ffi::F32> sin,
                       ffi::ResultBuffer<ffi::F32> logits, ffi::ResultBuffer<ffi::F32> probs, int64_t handle) {
  A2mHandle* h = as_handle(handle);
  if (!h || !is_windows(audio)) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "a2m_forward: audio must be [B, 2, 80000] f32");
  const int B = static_cast<int>(audio.dimensions()[0]);
  return status(h, a2m_forward(h, audio.typed_data(), B, cos.typed_data(), sin.typed_data(), static_cast<int>(cos.dimensions()[0]),
                               logits->typed_data(), probs->typed_data(), /*workspace_dev=*/nullptr, 0, stream),
                "a2m_forward");
}

// ---- a2m_forward_train: the dropout-enabled forward that records the tape (train.py:56-58).
// (params [n], audio, cos, sin; handle, dropout_rate, seed) -> logits.  `params` is the caller's parameter blob (leaves raveled in
// A2mLeafDesc order): copied into the handle's master weights and re-packed, so optax may own the parameters.
ffi::Error ForwardTrainImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> audio, ffi::Buffer<ffi::F32> cos,
                            ffi::Buffer<ffi::F32> sin, ffi::ResultBuffer<ffi::F32> logits, int64_t handle, float dropout_rate, int64_t seed) {
  A2mHandle* h = as_handle(handle);
  if (!h || !is_windows(audio)) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "a2m_forward_train: audio must be [B, 2, 80000] f32");
  if (static_cast<int64_t>(params.element_count()) != a2m_param_count(h))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "a2m_forward_train: params blob size != a2m_param_count (call a2m_train_init first)");
  int rc = a2m_set_params(h, params.typed_data(), stream);
  if (rc) return status(h, rc, "a2m_set_params");
  rc = a2m_set_dropout(h, dropout_rate, static_cast<uint64_t>(seed));
  if (rc) return status(h, rc, "a2m_set_dropout");
  const int B = static_cast<int>(audio.dimensions()[0]);
  return status(h, a2m_forward_train(h, audio.typed_data(), B, cos.typed_data(), sin.typed_data(), static_cast<int>(cos.dimensions()[0]),
                                     logits->typed_data(), nullptr, stream),
                "a2m_forward_train");
}

// ---- a2m_backward_dlogits: the custom_vjp backward.  (dlogits [B,250,90]; handle) -> grads [n] = J^T dlogits of the last forward_train.
ffi::Error BackwardImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> dlogits, ffi::ResultBuffer<ffi::F32> grads, int64_t handle) {
  A2mHandle* h = as_handle(handle);
  if (!h) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "a2m_backward: null handle");
  if (static_cast<int64_t>(grads->element_count()) != a2m_param_count(h))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "a2m_backward: grads blob size != a2m_param_count");
  if (cudaMemsetAsync(grads->typed_data(), 0, grads->size_bytes(), stream) != cudaSuccess)      // the C entry point accumulates
    return ffi::Error(ffi::ErrorCode::kInternal, "a2m_backward: cudaMemsetAsync failed");
  return status(h, a2m_backward_dlogits(h, dlogits.typed_data(), grads->typed_data(), stream), "a2m_backward_dlogits");
}

// ---- a2m_loss_and_grad: compute_loss under value_and_grad in ONE call (train.py:48-62).
// (params [n], audio, labels [B,250,90], cos, sin; handle, scale, dropout_rate, seed) -> (loss [1], grads [n])
ffi::Error LossAndGradImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> audio, ffi::Buffer<ffi::F32> labels,
                           ffi::Buffer<ffi::F32> cos, ffi::Buffer<ffi::F32> sin, ffi::ResultBuffer<ffi::F32> loss,
                           ffi::ResultBuffer<ffi::F32> grads, int64_t handle, float scale, float dropout_rate, int64_t seed) {
  A2mHandle* h = as_handle(handle);
  if (!h || !is_windows(audio)) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "a2m_loss_and_grad: audio must be [B, 2, 80000] f32");
  if (static_cast<int64_t>(params.element_count()) != a2m_param_count(h) || grads->element_count() != params.element_count())
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "a2m_loss_and_grad: params / grads blob size != a2m_param_count");
  const int B = static_cast<int>(audio.dimensions()[0]);
  int rc = a2m_set_params(h, params.typed_data(), stream);
  if (rc) return status(h, rc, "a2m_set_params");
  rc = a2m_set_dropout(h, dropout_rate, static_cast<uint64_t>(seed));
  if (rc) return status(h, rc, "a2m_set_dropout");
  rc = a2m_forward_train(h, audio.typed_data(), B, cos.typed_data(), sin.typed_data(), static_cast<int>(cos.dimensions()[0]), nullptr, nullptr,
                         stream);
  if (rc) return status(h, rc, "a2m_forward_train");
  if (cudaMemsetAsync(grads->typed_data(), 0, grads->size_bytes(), stream) != cudaSuccess ||
      cudaMemsetAsync(loss->typed_data(), 0, sizeof(float), stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "a2m_loss_and_grad: cudaMemsetAsync failed");
  return status(h, a2m_backward(h, labels.typed_data(), scale, grads->typed_data(), loss->typed_data(), stream), "a2m_backward");
}

// ---- a2m_allreduce: data-parallel mean of the gradient blob of the last backward, in place (the handle's own communicator,
// a2m_comm_init; bucket 0 overlapped with the tail of the backward).  (grads [n]) -> grads [n], aliased by the caller
// (input_output_aliases={0: 0}); the call is an ordering point only, the reduction runs on the blob a2m_backward wrote.
ffi::Error AllReduceImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> grads_in, ffi::ResultBuffer<ffi::F32> grads_out, int64_t handle) {
  A2mHandle* h = as_handle(handle);
  if (!h) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "a2m_allreduce: null handle");
  if (grads_in.typed_data() != grads_out->typed_data())
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "a2m_allreduce: bind with input_output_aliases={0: 0}");
  return status(h, a2m_allreduce_grads(h, nullptr, stream), "a2m_allreduce_grads");
}

// ---- a2m_adamw: optax.adamw + clip_by_global_norm on the updates, applied to the handle's master weights (train.py:324-325).
// (grads [n]; handle, lr, b1, b2, eps, weight_decay, grad_divisor, clip_norm, step) -> (params [n], stats [2])
ffi::Error AdamWImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> grads, ffi::ResultBuffer<ffi::F32> params, ffi::ResultBuffer<ffi::F32> stats,
                     int64_t handle, float lr, float b1, float b2, float eps, float weight_decay, float grad_divisor, float clip_norm,
                     int32_t step) {
  A2mHandle* h = as_handle(handle);
  if (!h) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "a2m_adamw: null handle");
  if (static_cast<int64_t>(grads.element_count()) != a2m_param_count(h) || params->element_count() != grads.element_count() ||
      stats->element_count() != 2)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "a2m_adamw: blob sizes");
  int rc = a2m_adamw_step(h, grads.typed_data(), lr, b1, b2, eps, weight_decay, grad_divisor, clip_norm, step, stats->typed_data(), stream);
  if (rc) return status(h, rc, "a2m_adamw_step");
  return status(h, a2m_get_params(h, params->typed_data(), stream), "a2m_get_params");
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(A2mForward, ForwardImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()   // audio
                                  .Arg<ffi::Buffer<ffi::F32>>()   // rope cos
                                  .Arg<ffi::Buffer<ffi::F32>>()   // rope sin
                                  .Ret<ffi::Buffer<ffi::F32>>()   // logits
                                  .Ret<ffi::Buffer<ffi::F32>>()   // probs
                                  .Attr<int64_t>("handle"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(A2mForwardTrain, ForwardTrainImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()   // params blob
                                  .Arg<ffi::Buffer<ffi::F32>>()   // audio
                                  .Arg<ffi::Buffer<ffi::F32>>()   // rope cos
                                  .Arg<ffi::Buffer<ffi::F32>>()   // rope sin
                                  .Ret<ffi::Buffer<ffi::F32>>()   // logits
                                  .Attr<int64_t>("handle")
                                  .Attr<float>("dropout_rate")
                                  .Attr<int64_t>("seed"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(A2mBackward, BackwardImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()   // dlogits
                                  .Ret<ffi::Buffer<ffi::F32>>()   // grads blob
                                  .Attr<int64_t>("handle"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(A2mLossAndGrad, LossAndGradImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()   // params blob
                                  .Arg<ffi::Buffer<ffi::F32>>()   // audio
                                  .Arg<ffi::Buffer<ffi::F32>>()   // labels
                                  .Arg<ffi::Buffer<ffi::F32>>()   // rope cos
                                  .Arg<ffi::Buffer<ffi::F32>>()   // rope sin
                                  .Ret<ffi::Buffer<ffi::F32>>()   // loss [1]
                                  .Ret<ffi::Buffer<ffi::F32>>()   // grads blob
                                  .Attr<int64_t>("handle")
                                  .Attr<float>("scale")
                                  .Attr<float>("dropout_rate")
                                  .Attr<int64_t>("seed"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(A2mAllReduce, AllReduceImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()   // grads blob (aliased to the result)
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int64_t>("handle"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(A2mAdamW, AdamWImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()   // grads blob
                                  .Ret<ffi::Buffer<ffi::F32>>()   // new params blob
                                  .Ret<ffi::Buffer<ffi::F32>>()   // stats [2]
                                  .Attr<int64_t>("handle")
                                  .Attr<float>("lr")
                                  .Attr<float>("b1")
                                  .Attr<float>("b2")
                                  .Attr<float>("eps")
                                  .Attr<float>("weight_decay")
                                  .Attr<float>("grad_divisor")
                                  .Attr<float>("clip_norm")
                                  .Attr<int32_t>("step"));
