// Lane types of the step kernel: a thread advances ONE vehicle in plain `float` registers or TWO vehicles (the same
// slot of two environments, hence the same airframe type) in packed `f2` registers.
//
// f2 is a 64-bit register pair driven by the sm_100 packed FP32 instructions (PTX fma/mul/add .f32x2 -> SASS FFMA2 /
// FMUL2 / FADD2): one issue slot performs the operation for both vehicles.  The step kernel is bound by instruction
// issue (profiles/r01_*), two thirds of its instructions are FP32 FMA-pipe operations, and FFMA2 occupies the FMA pipe
// exactly as long as two FFMA would - so packing two vehicles per thread halves the issue slots of the arithmetic
// without adding pipe cycles.  ptxas treats `mov.b64 {lo, hi}` as register aliasing (no MOV is emitted), folds
// negation and |x| into the packed operands, and turns a pair built from one scalar into a broadcast operand
// (`R6.F32`), so per-type constants loaded once from shared memory serve both vehicles of the thread.
// MUFU, min / max, compares and selects have no packed form: they run once per half on the aliased registers.
//
// Both halves are computed with the same IEEE round-to-nearest operations in the same order, so a vehicle's result does
// not depend on which half (or which thread) it occupies.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct f2 { unsigned long long v; };
struct m2 { bool a, b; };   // per-half predicate

template <class T> struct Lanes;
template <> struct Lanes<float> { static constexpr int N = 1; typedef bool Mask; };
template <> struct Lanes<f2>    { static constexpr int N = 2; typedef m2 Mask; };

// ---- pack / unpack ------------------------------------------------------------------------------------------------
__device__ __forceinline__ f2 f2_make(float a, float b) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void f2_split(f2 p, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v)); }
__device__ __forceinline__ float f2_lo(f2 p) { float a, b; f2_split(p, a, b); return a; }
__device__ __forceinline__ float f2_hi(f2 p) { float a, b; f2_split(p, a, b); return b; }
__device__ __forceinline__ f2 ld2(const float2& c) { return f2_make(c.x, c.y); }   // a float2 from memory as a register pair
__device__ __forceinline__ f2 to2(f2 x) { return x; }
__device__ __forceinline__ f2 to2(float x) { return f2_make(x, x); }

template <class T> __device__ __forceinline__ T ds_bc(float c);                       // broadcast
template <> __device__ __forceinline__ float ds_bc<float>(float c) { return c; }
template <> __device__ __forceinline__ f2 ds_bc<f2>(float c) { return f2_make(c, c); }

template <class T> __device__ __forceinline__ T ds_pack(const float* x);              // x[Lanes<T>::N]
template <> __device__ __forceinline__ float ds_pack<float>(const float* x) { return x[0]; }
template <> __device__ __forceinline__ f2 ds_pack<f2>(const float* x) { return f2_make(x[0], x[1]); }

__device__ __forceinline__ float ds_lane(float x, int) { return x; }
__device__ __forceinline__ float ds_lane(f2 x, int h) { float a, b; f2_split(x, a, b); return h ? b : a; }

// ---- packed arithmetic (FFMA2 / FMUL2 / FADD2; .ftz like the scalar code, which is compiled with -ftz=true) -------
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) { f2 r; asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) { f2 r; asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 f2_sub(f2 a, f2 b) { f2 r; asm("sub.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }

__device__ __forceinline__ f2 operator+(f2 a, f2 b) { return f2_add(a, b); }
__device__ __forceinline__ f2 operator+(f2 a, float b) { return f2_add(a, to2(b)); }
__device__ __forceinline__ f2 operator+(float a, f2 b) { return f2_add(to2(a), b); }
__device__ __forceinline__ f2 operator-(f2 a, f2 b) { return f2_sub(a, b); }
__device__ __forceinline__ f2 operator-(f2 a, float b) { return f2_sub(a, to2(b)); }
__device__ __forceinline__ f2 operator-(float a, f2 b) { return f2_sub(to2(a), b); }
__device__ __forceinline__ f2 operator*(f2 a, f2 b) { return f2_mul(a, b); }
__device__ __forceinline__ f2 operator*(f2 a, float b) { return f2_mul(a, to2(b)); }
__device__ __forceinline__ f2 operator*(float a, f2 b) { return f2_mul(to2(a), b); }
// Sign flip / magnitude WITHOUT flush-to-zero semantics: under -ftz=true the compiler's own `-x` / fabsf(x) are neg.ftz /
// abs.ftz, which ptxas must materialise (FADD.FTZ -x, -RZ); the plain forms fold into the consumer's operand modifier.
__device__ __forceinline__ float ds_neg(float x) { float r; asm("neg.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ds_abs(float x) { float r; asm("abs.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ f2 ds_neg(f2 a) { float x, y; f2_split(a, x, y); return f2_make(ds_neg(x), ds_neg(y)); }
__device__ __forceinline__ f2 operator-(f2 a) { return ds_neg(a); }
__device__ __forceinline__ f2& operator+=(f2& a, f2 b) { a = f2_add(a, b); return a; }
__device__ __forceinline__ f2& operator-=(f2& a, f2 b) { a = f2_sub(a, b); return a; }

// fused multiply-add, any mix of float / f2 operands: all-float resolves to the non-template overload (FFMA), anything
// else to the packed one (FFMA2 with broadcast operands).  The kernels spell every fusion out through this function:
// packed inline PTX is not contracted by the compiler, and the explicit form is what raised the FFMA share of the FP32
// instructions (the r01 kernel left 34 % FMUL + 18 % FADD unfused).
__device__ __forceinline__ float ds_fma(float a, float b, float c) { return fmaf(a, b, c); }
template <class A, class B, class C>
__device__ __forceinline__ f2 ds_fma(A a, B b, C c) { return f2_fma(to2(a), to2(b), to2(c)); }

// ---- per-half operations ---------------------------------------------------------------------------------------------
// single-instruction MUFU forms (flush-to-zero, ~1 ulp): no denormal pre/post scaling around the SFU op
__device__ __forceinline__ float ds_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ds_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ds_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#define DS_F2_UNARY(name, expr)                                                    \
  __device__ __forceinline__ f2 name(f2 p) { float x, y; f2_split(p, x, y); float a = x; float ra = (expr); a = y; float rb = (expr); return f2_make(ra, rb); }
DS_F2_UNARY(ds_rcp, ds_rcp(a))
DS_F2_UNARY(ds_ex2, ds_ex2(a))
DS_F2_UNARY(ds_rsqrt, ds_rsqrt(a))
DS_F2_UNARY(ds_abs, ds_abs(a))
DS_F2_UNARY(ds_sqrt, sqrtf(a))
#undef DS_F2_UNARY
__device__ __forceinline__ float ds_sqrt(float x) { return sqrtf(x); }

__device__ __forceinline__ float ds_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ float ds_min(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ f2 ds_max(f2 p, float b) { float x, y; f2_split(p, x, y); return f2_make(fmaxf(x, b), fmaxf(y, b)); }
__device__ __forceinline__ f2 ds_min(f2 p, float b) { float x, y; f2_split(p, x, y); return f2_make(fminf(x, b), fminf(y, b)); }
__device__ __forceinline__ f2 ds_max(f2 p, f2 q) { float x, y, u, v; f2_split(p, x, y); f2_split(q, u, v); return f2_make(fmaxf(x, u), fmaxf(y, v)); }
__device__ __forceinline__ f2 ds_min(f2 p, f2 q) { float x, y, u, v; f2_split(p, x, y); f2_split(q, u, v); return f2_make(fminf(x, u), fminf(y, v)); }

// compares -> masks, selects
__device__ __forceinline__ bool ds_gt(float a, float b) { return a > b; }
__device__ __forceinline__ bool ds_lt(float a, float b) { return a < b; }
__device__ __forceinline__ bool ds_ge(float a, float b) { return a >= b; }
__device__ __forceinline__ m2 ds_gt(f2 p, float b) { float x, y; f2_split(p, x, y); return m2{x > b, y > b}; }
__device__ __forceinline__ m2 ds_lt(f2 p, float b) { float x, y; f2_split(p, x, y); return m2{x < b, y < b}; }
__device__ __forceinline__ m2 ds_ge(f2 p, float b) { float x, y; f2_split(p, x, y); return m2{x >= b, y >= b}; }
__device__ __forceinline__ bool ds_and(bool a, bool b) { return a && b; }
__device__ __forceinline__ m2 ds_and(m2 a, m2 b) { return m2{a.a && b.a, a.b && b.b}; }
__device__ __forceinline__ bool ds_any(bool a) { return a; }
__device__ __forceinline__ bool ds_any(m2 a) { return a.a || a.b; }
__device__ __forceinline__ float ds_sel(bool m, float a, float b) { return m ? a : b; }
__device__ __forceinline__ f2 ds_sel(m2 m, f2 p, f2 q) { float x, y, u, v; f2_split(p, x, y); f2_split(q, u, v); return f2_make(m.a ? x : u, m.b ? y : v); }
__device__ __forceinline__ f2 ds_sel(m2 m, f2 p, float q) { float x, y; f2_split(p, x, y); return f2_make(m.a ? x : q, m.b ? y : q); }
__device__ __forceinline__ f2 ds_sel(m2 m, float p, f2 q) { float u, v; f2_split(q, u, v); return f2_make(m.a ? p : u, m.b ? p : v); }
__device__ __forceinline__ f2 ds_sel(m2 m, float p, float q) { return f2_make(m.a ? p : q, m.b ? p : q); }

// 16-lane-wide indexed shuffle of every half
__device__ __forceinline__ float ds_shfl16(float x, int src) { return __shfl_sync(0xffffffffu, x, src, 16); }
__device__ __forceinline__ f2 ds_shfl16(f2 p, int src) {
  float x, y;
  f2_split(p, x, y);
  return f2_make(__shfl_sync(0xffffffffu, x, src, 16), __shfl_sync(0xffffffffu, y, src, 16));
}
