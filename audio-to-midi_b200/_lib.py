"""ctypes binding of include/a2m.h.  Fails loudly: there is no CPU or PyTorch fallback."""
from __future__ import annotations

import ctypes as C
import os

from .build import LIBS, build

_libs = {}


class A2mError(RuntimeError):
    pass


class LeafDesc(C.Structure):
    _fields_ = [("path", C.c_char_p), ("offset_bytes", C.c_uint64), ("ndim", C.c_int32), ("shape", C.c_int64 * 4)]


class MidiEvent(C.Structure):
    _fields_ = [("attack_time", C.c_uint64), ("note", C.c_uint8), ("duration", C.c_uint64), ("velocity", C.c_uint8)]


class MidiEventList(C.Structure):
    _fields_ = [("ptr", C.POINTER(MidiEvent)), ("length", C.c_size_t), ("_capacity", C.c_size_t)]


class StepProfile(C.Structure):
    _fields_ = [("kernel", C.c_char * 32), ("ms", C.c_float), ("flops", C.c_double), ("bytes", C.c_double)]


class MLMultiArrayWrapper3(C.Structure):
    _fields_ = [("strides", C.c_uint64 * 3), ("dims", C.c_uint64 * 3), ("data", C.c_void_p)]


EXPORTS = [
    "a2m_create", "a2m_destroy", "a2m_last_error", "a2m_load_weights", "a2m_workspace_bytes", "a2m_forward",
    "a2m_forward_host", "a2m_submit_host", "a2m_collect_host", "a2m_host_alloc", "a2m_host_free",
    "a2m_last_launch_count", "a2m_window_count", "a2m_prepare_windows", "a2m_window_losses", "a2m_profile_steps", "a2m_set_use_graph", "a2m_debug_read_timing", "a2m_debug_gemm_pair", "a2m_debug_forward_tap", "a2m_debug_gemm",
    "a2m_train_init", "a2m_set_dropout", "a2m_param_count", "a2m_get_params", "a2m_set_lr_multipliers", "a2m_forward_train", "a2m_backward",
    "a2m_grad_bucket_count", "a2m_grad_bucket_range", "a2m_stream_wait_grad_bucket",
    "a2m_adamw_step", "a2m_train_launch_count", "a2m_debug_wgrad", "a2m_profile_train_steps",
    "a2m_stitch_probs", "a2m_extract_events", "extract_midi_events", "free_midi_events", "a2m_to_frame_events",
    "a2m_create_ex", "a2m_submit_host_ex", "a2m_event_metrics", "a2m_set_params", "a2m_get_opt_state", "a2m_set_opt_state",
    "a2m_operand_format", "a2m_debug_round_operand", "a2m_stitch_probs_dev", "a2m_extract_events_dev", "a2m_backward_dlogits", "a2m_allreduce_grads", "a2m_comm_unique_id", "a2m_comm_init", "a2m_comm_get", "a2m_comm_destroy",
]


class A2mConfig(C.Structure):   # include/a2m.h
    _fields_ = [("device", C.c_int32), ("num_stages", C.c_int32), ("dims", C.c_int32 * 8), ("depths", C.c_int32 * 8),
                ("cnn_hidden_expansion_x2", C.c_int32), ("num_transformer_layers", C.c_int32), ("num_transformer_heads", C.c_int32),
                ("attention_size", C.c_int32), ("compressed_attention_kv_size", C.c_int32), ("transformer_intermediate", C.c_int32),
                ("use_graph", C.c_int32), ("use_pdl", C.c_int32)]


F32, F16 = 0, 1


def lib(precision: str = "bf16") -> C.CDLL:
    """Loads (building in-tree first if a compiler is present and the .so is stale) the C-ABI library in one of its two
    operand-format variants: "bf16" (training + inference) or "f16" (inference, binary16 tensor-core operands)."""
    if precision not in LIBS:
        raise ValueError(f"precision must be one of {sorted(LIBS)}, got {precision!r}")
    if precision in _libs:
        return _libs[precision]
    path = LIBS[precision]
    override = os.environ.get(f"A2M_LIB_{precision.upper()}")      # A/B experiments (tools/ab_build.sh): an alternative build of this variant
    if override:
        if not os.path.exists(override):
            raise FileNotFoundError(f"A2M_LIB_{precision.upper()}={override} does not exist")
        path = override
    elif not os.path.exists(path) or os.environ.get("A2M_REBUILD") == "1":
        path = build(precision=precision)
    else:
        try:
            path = build(precision=precision)  # no-op when up to date
        except RuntimeError:
            pass  # no nvcc on this box: use the prebuilt library that travelled with the tree
    L = C.CDLL(path)
    vp, i32, u32, u64, sz, f64 = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64, C.c_size_t, C.c_double
    L.a2m_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.a2m_create.restype = C.c_int
    L.a2m_destroy.argtypes = [vp]
    L.a2m_destroy.restype = None
    L.a2m_last_error.argtypes = [vp]
    L.a2m_last_error.restype = C.c_char_p
    L.a2m_load_weights.argtypes = [vp, vp, sz, C.POINTER(LeafDesc), i32]
    L.a2m_load_weights.restype = C.c_int
    L.a2m_workspace_bytes.argtypes = [vp, i32, i32]
    L.a2m_workspace_bytes.restype = sz
    L.a2m_forward.argtypes = [vp, vp, i32, vp, vp, i32, vp, vp, vp, sz, vp]
    L.a2m_forward.restype = C.c_int
    L.a2m_forward_host.argtypes = [vp, vp, i32, vp, vp, i32, vp, vp]
    L.a2m_forward_host.restype = C.c_int
    L.a2m_submit_host.argtypes = [vp, i32, vp, i32, vp, vp, i32, vp, vp]
    L.a2m_submit_host.restype = C.c_int
    L.a2m_collect_host.argtypes = [vp, i32]
    L.a2m_collect_host.restype = C.c_int
    L.a2m_host_alloc.argtypes = [sz]
    L.a2m_host_alloc.restype = vp
    L.a2m_host_free.argtypes = [vp]
    L.a2m_host_free.restype = None
    L.a2m_window_count.argtypes = [C.c_int64, f64]
    L.a2m_window_count.restype = C.c_int64
    L.a2m_prepare_windows.argtypes = [vp, vp, C.c_int64, f64, vp, C.c_int64, vp]
    L.a2m_prepare_windows.restype = C.c_int
    L.a2m_window_losses.argtypes = [vp, vp, vp, i32, vp, vp]
    L.a2m_window_losses.restype = C.c_int
    L.a2m_last_launch_count.argtypes = [vp]
    L.a2m_last_launch_count.restype = i32
    L.a2m_profile_steps.argtypes = [vp, i32, i32, i32, C.POINTER(StepProfile)]
    L.a2m_profile_steps.restype = i32
    L.a2m_set_use_graph.argtypes = [vp, i32]
    L.a2m_set_use_graph.restype = C.c_int
    L.a2m_debug_forward_tap.argtypes = [vp, vp, i32, vp, vp, i32, C.c_char_p, vp, sz, vp]
    L.a2m_debug_forward_tap.restype = C.c_int
    L.a2m_debug_gemm.argtypes = [vp, i32, i32, i32, i32, vp, i32, vp, u32, vp, vp, vp, vp, vp, vp]
    L.a2m_debug_gemm.restype = C.c_int
    f32 = C.c_float
    L.a2m_train_init.argtypes = [vp, vp, sz, C.POINTER(LeafDesc), i32]
    L.a2m_train_init.restype = C.c_int
    L.a2m_set_dropout.argtypes = [vp, f32, u64]
    L.a2m_set_dropout.restype = C.c_int
    L.a2m_param_count.argtypes = [vp]
    L.a2m_param_count.restype = C.c_int64
    L.a2m_get_params.argtypes = [vp, vp, vp]
    L.a2m_get_params.restype = C.c_int
    L.a2m_set_lr_multipliers.argtypes = [vp, vp, i32]
    L.a2m_set_lr_multipliers.restype = C.c_int
    L.a2m_forward_train.argtypes = [vp, vp, i32, vp, vp, i32, vp, vp, vp]
    L.a2m_forward_train.restype = C.c_int
    L.a2m_backward.argtypes = [vp, vp, f32, vp, vp, vp]
    L.a2m_backward.restype = C.c_int
    L.a2m_grad_bucket_count.argtypes = [vp]
    L.a2m_grad_bucket_count.restype = i32
    L.a2m_grad_bucket_range.argtypes = [vp, i32, C.POINTER(sz), C.POINTER(sz)]
    L.a2m_grad_bucket_range.restype = C.c_int
    L.a2m_stream_wait_grad_bucket.argtypes = [vp, i32, vp]
    L.a2m_stream_wait_grad_bucket.restype = C.c_int
    L.a2m_adamw_step.argtypes = [vp, vp, f32, f32, f32, f32, f32, f32, f32, i32, vp, vp]
    L.a2m_adamw_step.restype = C.c_int
    L.a2m_train_launch_count.argtypes = [vp]
    L.a2m_train_launch_count.restype = i32
    L.a2m_profile_train_steps.argtypes = [vp, i32, i32, i32, C.POINTER(StepProfile)]
    L.a2m_profile_train_steps.restype = i32
    L.a2m_debug_wgrad.argtypes = [vp, i32, i32, i32, vp, i32, vp, i32, vp, vp]
    L.a2m_debug_wgrad.restype = C.c_int
    L.a2m_stitch_probs.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64, f64, f64, vp]
    L.a2m_stitch_probs.restype = C.c_int64
    L.a2m_extract_events.argtypes = [vp, C.c_int64, C.c_int64]
    L.a2m_extract_events.restype = C.POINTER(MidiEventList)
    L.extract_midi_events.argtypes = [MLMultiArrayWrapper3, f64, f64]
    L.extract_midi_events.restype = C.POINTER(MidiEventList)
    L.free_midi_events.argtypes = [C.POINTER(MidiEventList)]
    L.free_midi_events.restype = None
    L.a2m_to_frame_events.argtypes = [C.POINTER(MidiEvent), C.c_int64, C.c_int64, vp]
    L.a2m_to_frame_events.restype = C.c_int
    L.a2m_create_ex.argtypes = [C.POINTER(A2mConfig), C.POINTER(vp)]
    L.a2m_create_ex.restype = C.c_int
    L.a2m_submit_host_ex.argtypes = [vp, i32, vp, i32, i32, vp, vp, i32, vp, vp, i32]
    L.a2m_submit_host_ex.restype = C.c_int
    L.a2m_event_metrics.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp, vp]
    L.a2m_event_metrics.restype = C.c_int
    L.a2m_set_params.argtypes = [vp, vp, vp]
    L.a2m_set_params.restype = C.c_int
    L.a2m_get_opt_state.argtypes = [vp, vp, vp, vp]
    L.a2m_get_opt_state.restype = C.c_int
    L.a2m_set_opt_state.argtypes = [vp, vp, vp, vp]
    L.a2m_set_opt_state.restype = C.c_int
    L.a2m_backward_dlogits.argtypes = [vp, vp, vp, vp]
    L.a2m_backward_dlogits.restype = C.c_int
    L.a2m_allreduce_grads.argtypes = [vp, vp, vp]
    L.a2m_allreduce_grads.restype = C.c_int
    L.a2m_comm_unique_id.argtypes = [vp]
    L.a2m_comm_unique_id.restype = C.c_int
    L.a2m_comm_init.argtypes = [vp, vp, i32, i32]
    L.a2m_comm_init.restype = C.c_int
    L.a2m_comm_get.argtypes = [vp]
    L.a2m_comm_get.restype = vp
    L.a2m_comm_destroy.argtypes = [vp]
    L.a2m_comm_destroy.restype = C.c_int
    L.a2m_stitch_probs_dev.argtypes = [vp, vp, C.c_int64, C.c_int64, C.c_int64, f64, f64, vp, C.c_int64, vp]
    L.a2m_stitch_probs_dev.restype = C.c_int64
    L.a2m_extract_events_dev.argtypes = [vp, vp, C.c_int64, C.c_int64, vp, C.c_int64, vp, vp]
    L.a2m_extract_events_dev.restype = C.c_int
    L.a2m_operand_format.argtypes = []
    L.a2m_operand_format.restype = C.c_char_p
    L.a2m_debug_round_operand.argtypes = [vp, vp, C.c_int64]
    L.a2m_debug_round_operand.restype = C.c_int
    if L.a2m_operand_format().decode() != precision:
        raise A2mError(f"{path} reports operand format {L.a2m_operand_format().decode()!r}, expected {precision!r}")
    _libs[precision] = L
    return L


def check(handle, rc: int, what: str, L=None):
    if rc != 0:
        msg = (L or lib()).a2m_last_error(handle)
        raise A2mError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")
