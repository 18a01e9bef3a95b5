/* TEST STUB -- not XLA.  Just enough of xla/ffi/api/c_api.h for `g++ -fsyntax-only` of csrc/a2m_xla_ffi.cc in an image that
 * has no jaxlib.  See ffi.h next to it. */
#pragma once
struct XLA_FFI_Error;
struct XLA_FFI_CallFrame;
