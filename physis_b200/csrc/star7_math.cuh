// Arithmetic of the 7-point diffusion update shared by the single-sweep kernels
// (star7.cu) and the fused two-sweep kernel (star7_pair.cu): the reference's
// left-to-right order with separately rounded products and sums
// (examples/diffusion-benchmark/diffusion3d_physis.c:55-56 as the REFERENCE target
// evaluates it), so every form is bit-identical to the oracle.  `A` is any struct
// with the coefficient members cc, cw, ce, cs, cn, cb, ct.
#pragma once
#include "sweep_common.cuh"

namespace physis_b200 {
namespace sweep {

template <typename T, typename A>
__device__ __forceinline__ T Point7(const A &a, T c, T w, T e, T s, T n, T b, T t) {
  // ((((((cc*c + cw*w) + ce*e) + cs*s) + cn*n) + cb*b) + ct*t)
  T r = MulRn(a.cc, c);
  r = AddRn(r, MulRn(a.cw, w));
  r = AddRn(r, MulRn(a.ce, e));
  r = AddRn(r, MulRn(a.cs, s));
  r = AddRn(r, MulRn(a.cn, n));
  r = AddRn(r, MulRn(a.cb, b));
  r = AddRn(r, MulRn(a.ct, t));
  return r;
}

namespace v2 {

typedef unsigned long long u64;

__device__ __forceinline__ u64 Pack(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void Unpack(u64 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 Add2(u64 a, u64 b) {
  u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// out vector from centre c, neighbours: west/east scalars of the vector's ends,
// and the s, n, b, t vectors.  ((((((cc*c + cw*w) + ce*e) + cs*s) + cn*n) + cb*b) + ct*t)
template <int FP, typename A>
__device__ __forceinline__ float4 Vec7(const A &a, float4 c, float wv, float ev,
                                       float4 s, float4 n, float4 b, float4 t) {
  float4 o;
  if (FP == 0) {
    o.x = Point7<float>(a, c.x, wv, c.y, s.x, n.x, b.x, t.x);
    o.y = Point7<float>(a, c.y, c.x, c.z, s.y, n.y, b.y, t.y);
    o.z = Point7<float>(a, c.z, c.y, c.w, s.z, n.z, b.z, t.z);
    o.w = Point7<float>(a, c.w, c.z, ev, s.w, n.w, b.w, t.w);
  } else {
    // separately rounded products, then packed adds in the reference's order
    u64 r01 = Pack(MulRn(a.cc, c.x), MulRn(a.cc, c.y));
    u64 r23 = Pack(MulRn(a.cc, c.z), MulRn(a.cc, c.w));
    r01 = Add2(r01, Pack(MulRn(a.cw, wv), MulRn(a.cw, c.x)));
    r23 = Add2(r23, Pack(MulRn(a.cw, c.y), MulRn(a.cw, c.z)));
    r01 = Add2(r01, Pack(MulRn(a.ce, c.y), MulRn(a.ce, c.z)));
    r23 = Add2(r23, Pack(MulRn(a.ce, c.w), MulRn(a.ce, ev)));
    r01 = Add2(r01, Pack(MulRn(a.cs, s.x), MulRn(a.cs, s.y)));
    r23 = Add2(r23, Pack(MulRn(a.cs, s.z), MulRn(a.cs, s.w)));
    r01 = Add2(r01, Pack(MulRn(a.cn, n.x), MulRn(a.cn, n.y)));
    r23 = Add2(r23, Pack(MulRn(a.cn, n.z), MulRn(a.cn, n.w)));
    r01 = Add2(r01, Pack(MulRn(a.cb, b.x), MulRn(a.cb, b.y)));
    r23 = Add2(r23, Pack(MulRn(a.cb, b.z), MulRn(a.cb, b.w)));
    r01 = Add2(r01, Pack(MulRn(a.ct, t.x), MulRn(a.ct, t.y)));
    r23 = Add2(r23, Pack(MulRn(a.ct, t.z), MulRn(a.ct, t.w)));
    Unpack(r01, o.x, o.y);
    Unpack(r23, o.z, o.w);
  }
  return o;
}
template <int FP, typename A>
__device__ __forceinline__ double2 Vec7(const A &a, double2 c, double wv, double ev,
                                        double2 s, double2 n, double2 b, double2 t) {
  double2 o;
  o.x = Point7<double>(a, c.x, wv, c.y, s.x, n.x, b.x, t.x);
  o.y = Point7<double>(a, c.y, c.x, ev, s.y, n.y, b.y, t.y);
  return o;
}

// ---- equal neighbour coefficients (cw == ce == cs == cn == cb == ct =: c6, bit for bit) ----
// The product c6*v of a value v is then the same number wherever v is a neighbour, so a
// sweep can form it once per point and reuse it for all six roles: 2 multiplies per point
// (cc*v and c6*v) instead of 7.  The additions keep the reference's order, so the result
// is still bit-identical.  Scale() forms the products of a vector, Sum7() adds them up:
//   q = cc*c, wP / eP = products of the x neighbours of the vector's ends, pc = c6*c (its
//   elements are the x neighbours of each other), ps / pn / pb / pt = products of the s, n,
//   b, t vectors.
__device__ __forceinline__ float4 Scale(float k, const float4 &v) {
  return make_float4(MulRn(k, v.x), MulRn(k, v.y), MulRn(k, v.z), MulRn(k, v.w));
}
__device__ __forceinline__ double2 Scale(double k, const double2 &v) {
  return make_double2(MulRn(k, v.x), MulRn(k, v.y));
}
template <int FP>
__device__ __forceinline__ float4 Sum7(float4 q, float wP, float eP, float4 pc, float4 ps,
                                       float4 pn, float4 pb, float4 pt) {
  // ((q + w) + e): the x terms pair up across vector elements, so they stay scalar
  const float r0 = AddRn(AddRn(q.x, wP), pc.y);
  const float r1 = AddRn(AddRn(q.y, pc.x), pc.z);
  const float r2 = AddRn(AddRn(q.z, pc.y), pc.w);
  const float r3 = AddRn(AddRn(q.w, pc.z), eP);
  float4 o;
  if (FP == 0) {
    o.x = AddRn(AddRn(AddRn(AddRn(r0, ps.x), pn.x), pb.x), pt.x);
    o.y = AddRn(AddRn(AddRn(AddRn(r1, ps.y), pn.y), pb.y), pt.y);
    o.z = AddRn(AddRn(AddRn(AddRn(r2, ps.z), pn.z), pb.z), pt.z);
    o.w = AddRn(AddRn(AddRn(AddRn(r3, ps.w), pn.w), pb.w), pt.w);
  } else {
    u64 r01 = Pack(r0, r1), r23 = Pack(r2, r3);
    r01 = Add2(r01, Pack(ps.x, ps.y)); r23 = Add2(r23, Pack(ps.z, ps.w));
    r01 = Add2(r01, Pack(pn.x, pn.y)); r23 = Add2(r23, Pack(pn.z, pn.w));
    r01 = Add2(r01, Pack(pb.x, pb.y)); r23 = Add2(r23, Pack(pb.z, pb.w));
    r01 = Add2(r01, Pack(pt.x, pt.y)); r23 = Add2(r23, Pack(pt.z, pt.w));
    Unpack(r01, o.x, o.y);
    Unpack(r23, o.z, o.w);
  }
  return o;
}
template <int FP>
__device__ __forceinline__ double2 Sum7(double2 q, double wP, double eP, double2 pc, double2 ps,
                                        double2 pn, double2 pb, double2 pt) {
  double2 o;
  o.x = AddRn(AddRn(AddRn(AddRn(AddRn(AddRn(q.x, wP), pc.y), ps.x), pn.x), pb.x), pt.x);
  o.y = AddRn(AddRn(AddRn(AddRn(AddRn(AddRn(q.y, pc.x), eP), ps.y), pn.y), pb.y), pt.y);
  return o;
}

__device__ __forceinline__ float First(const float4 &v) { return v.x; }
__device__ __forceinline__ float Last(const float4 &v) { return v.w; }
__device__ __forceinline__ double First(const double2 &v) { return v.x; }
__device__ __forceinline__ double Last(const double2 &v) { return v.y; }

}  // namespace v2

}  // namespace sweep
}  // namespace physis_b200
