// Packed fp32 arithmetic of sm_100 (two IEEE fp32 operations per issued instruction), shared by the tensor-core
// kernels' CUDA-core roles and the fp32 kernels.  Part of libfsvc.so (device code only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fsvc {

// ---- packed fp32 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2 do two IEEE fp32 operations per issued instruction).
// The transform and epilogue roles are bound by instruction issue (4 warps per scheduler, ~0.55 issue slots per cycle in
// the last stage), so every pair of scalar operations folded into one packed instruction is time, at identical results.
__device__ __forceinline__ void fma2(float& x0, float& x1, float a0, float a1, float c0, float c1) {  // x = x * a + c
  asm("{\n.reg .b64 x, a, c;\nmov.b64 x, {%0,%1};\nmov.b64 a, {%2,%3};\nmov.b64 c, {%4,%5};\n"
      "fma.rn.f32x2 x, x, a, c;\nmov.b64 {%0,%1}, x;\n}"
      : "+f"(x0), "+f"(x1) : "f"(a0), "f"(a1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void add2(float& x0, float& x1, float a0, float a1) {  // x += a
  asm("{\n.reg .b64 x, a;\nmov.b64 x, {%0,%1};\nmov.b64 a, {%2,%3};\nadd.rn.f32x2 x, x, a;\nmov.b64 {%0,%1}, x;\n}"
      : "+f"(x0), "+f"(x1) : "f"(a0), "f"(a1));
}
__device__ __forceinline__ void sub2(float& d0, float& d1, float x0, float x1, float a0, float a1) {  // d = x - a
  asm("{\n.reg .b64 x, a;\nmov.b64 x, {%2,%3};\nmov.b64 a, {%4,%5};\nsub.rn.f32x2 x, x, a;\nmov.b64 {%0,%1}, x;\n}"
      : "=f"(d0), "=f"(d1) : "f"(x0), "f"(x1), "f"(a0), "f"(a1));
}
__device__ __forceinline__ void mul2(float& d0, float& d1, float x0, float x1, float s) {  // d = x * s
  asm("{\n.reg .b64 x, a;\nmov.b64 x, {%2,%3};\nmov.b64 a, {%4,%4};\nmul.rn.f32x2 x, x, a;\nmov.b64 {%0,%1}, x;\n}"
      : "=f"(d0), "=f"(d1) : "f"(x0), "f"(x1), "f"(s));
}
__device__ __forceinline__ void fma2_acc(float& c0, float& c1, float a0, float a1, float b0, float b1) {  // c += a * b
  asm("{\n.reg .b64 x, a, c;\nmov.b64 x, {%2,%3};\nmov.b64 a, {%4,%5};\nmov.b64 c, {%0,%1};\n"
      "fma.rn.f32x2 c, x, a, c;\nmov.b64 {%0,%1}, c;\n}"
      : "+f"(c0), "+f"(c1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
}  // namespace fsvc
