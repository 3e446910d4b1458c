// diffgeo.cu - per-element gradient and integrated divergence.
//
// Replaces lapy/diffgeo.py:222-300 (tria_compute_gradient), :303-387 (tria_compute_divergence),
// :846-922 (tet_compute_gradient) and :925-1006 (tet_compute_divergence).  Geometry (edges,
// normals, cotangents, volumes) is evaluated in the dtype of the caller's vertices with unfused
// IEEE ops, the function-dependent part in fp64 - the promotion NumPy performs.  The vertex
// scatter of the divergence is a gather over the vertex->element incidence in a fixed order, so
// results are bit-reproducible (the reference's order is SciPy's; values agree to rounding).
#include "common.cuh"

namespace lb {

void ensure_incidence(lb_mesh *mesh);  // assembly.cu

using ED = Ex<double>;

__device__ __forceinline__ Vec3<double> scale3(double s, const Vec3<double> &a) {
    return {ED::mul(s, a.x), ED::mul(s, a.y), ED::mul(s, a.z)};
}
__device__ __forceinline__ Vec3<double> add3(const Vec3<double> &a, const Vec3<double> &b) {
    return {ED::add(a.x, b.x), ED::add(a.y, b.y), ED::add(a.z, b.z)};
}

template <class T>
__global__ void __launch_bounds__(256) tria_gradient_kernel(const typename Ex<T>::V4 *__restrict__ v4,
                                                            const int4 *__restrict__ t4, int64_t nt,
                                                            const double *__restrict__ f, int nf,
                                                            double *__restrict__ g) {
    using E = Ex<T>;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nt) return;
    const int4 ti = __ldg(t4 + e);
    const Vec3<T> p0 = load_vertex<T>(v4, ti.x), p1 = load_vertex<T>(v4, ti.y), p2 = load_vertex<T>(v4, ti.z);
    const Vec3<T> e2 = vsub(p1, p0), e0 = vsub(p2, p1), e1 = vsub(p0, p2);
    Vec3<T> nrm = vcross(e2, vneg(e1));
    T ln = E::sqrt(vdot(nrm, nrm));
    if (ln < E::eps()) ln = (T)1;
    const T lni = E::div((T)1, ln);
    nrm = {E::mul(nrm.x, lni), E::mul(nrm.y, lni), E::mul(nrm.z, lni)};
    const Vec3<double> nd = vwiden(nrm), d0 = vwiden(e0), d1 = vwiden(e1), d2 = vwiden(e2);
    const double lnid = (double)lni;
    for (int k = 0; k < nf; k++) {
        const double f0 = f[(int64_t)ti.x * nf + k], f1 = f[(int64_t)ti.y * nf + k], f2 = f[(int64_t)ti.z * nf + k];
        const Vec3<double> s = add3(add3(scale3(f0, d0), scale3(f1, d1)), scale3(f2, d2));
        const Vec3<double> cr = vcross(nd, s);
        double *o = g + (e * nf + k) * 3;
        o[0] = ED::mul(lnid, cr.x);
        o[1] = ED::mul(lnid, cr.y);
        o[2] = ED::mul(lnid, cr.z);
    }
}

template <class T>
__global__ void __launch_bounds__(256) tet_gradient_kernel(const typename Ex<T>::V4 *__restrict__ v4,
                                                           const int4 *__restrict__ t4, int64_t nt,
                                                           const double *__restrict__ f, int nf,
                                                           double *__restrict__ g) {
    using E = Ex<T>;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nt) return;
    const int4 ti = __ldg(t4 + e);
    const Vec3<T> p0 = load_vertex<T>(v4, ti.x), p1 = load_vertex<T>(v4, ti.y);
    const Vec3<T> p2 = load_vertex<T>(v4, ti.z), p3 = load_vertex<T>(v4, ti.w);
    const Vec3<T> e0 = vsub(p1, p0), e2 = vsub(p0, p2), e3 = vsub(p3, p0), e4 = vsub(p3, p1), e5 = vsub(p3, p2);
    T vol = fabs(vdot(e3, vcross(e0, e2)));
    if (vol < E::eps()) vol = (T)1;
    const double voli = (double)E::div((T)1, vol);
    const Vec3<double> g1 = vwiden(vcross(e2, e5)), g2 = vwiden(vcross(e3, e4)), g3 = vwiden(vcross(vneg(e2), e0));
    for (int k = 0; k < nf; k++) {
        const double f0 = f[(int64_t)ti.x * nf + k];
        const double d1 = ED::sub(f[(int64_t)ti.y * nf + k], f0), d2 = ED::sub(f[(int64_t)ti.z * nf + k], f0);
        const double d3 = ED::sub(f[(int64_t)ti.w * nf + k], f0);
        const Vec3<double> s = add3(add3(scale3(d1, g1), scale3(d2, g2)), scale3(d3, g3));
        double *o = g + (e * nf + k) * 3;
        o[0] = ED::mul(voli, s.x);
        o[1] = ED::mul(voli, s.y);
        o[2] = ED::mul(voli, s.z);
    }
}

__device__ __forceinline__ double dot_wide(const Vec3<double> &a, const double *x) {
    return ED::add(ED::add(ED::mul(a.x, x[0]), ED::mul(a.y, x[1])), ED::mul(a.z, x[2]));
}

// per element-corner contributions: cx[(e*K + corner)*nf + k]
template <class T>
__global__ void __launch_bounds__(256) tria_div_corner_kernel(const typename Ex<T>::V4 *__restrict__ v4,
                                                              const int4 *__restrict__ t4, int64_t nt,
                                                              const double *__restrict__ x, int nf,
                                                              double *__restrict__ cx) {
    using E = Ex<T>;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nt) return;
    const int4 ti = __ldg(t4 + e);
    const Vec3<T> p0 = load_vertex<T>(v4, ti.x), p1 = load_vertex<T>(v4, ti.y), p2 = load_vertex<T>(v4, ti.z);
    const Vec3<T> e2 = vsub(p1, p0), e0 = vsub(p2, p1), e1 = vsub(p0, p2);
    const Vec3<T> nrm = vcross(e2, vneg(e1));
    T ln = E::sqrt(vdot(nrm, nrm));
    if (ln < E::eps()) ln = (T)1;
    const T cot0 = E::div(vdot(e2, vneg(e1)), ln), cot1 = E::div(vdot(e0, vneg(e2)), ln);
    const T cot2 = E::div(vdot(e1, vneg(e0)), ln);
    const Vec3<T> c0 = {E::mul(cot0, e0.x), E::mul(cot0, e0.y), E::mul(cot0, e0.z)};
    const Vec3<T> c1 = {E::mul(cot1, e1.x), E::mul(cot1, e1.y), E::mul(cot1, e1.z)};
    const Vec3<T> c2 = {E::mul(cot2, e2.x), E::mul(cot2, e2.y), E::mul(cot2, e2.z)};
    const Vec3<double> w0 = vwiden(vsub(c2, c1)), w1 = vwiden(vsub(c0, c2)), w2 = vwiden(vsub(c1, c0));
    for (int k = 0; k < nf; k++) {
        const double *xv = x + (e * nf + k) * 3;
        cx[(e * 3 + 0) * nf + k] = dot_wide(w0, xv);
        cx[(e * 3 + 1) * nf + k] = dot_wide(w1, xv);
        cx[(e * 3 + 2) * nf + k] = dot_wide(w2, xv);
    }
}

// flux form (tria_compute_divergence2, lapy/diffgeo.py:390-469): x_k = <cross(e_k, n), X> with the
// unit normal n = cross(e2, -e1) / |.| (eps clamp -> 1), geometry in the dtype of the vertices
template <class T>
__global__ void __launch_bounds__(256) tria_div2_corner_kernel(const typename Ex<T>::V4 *__restrict__ v4,
                                                               const int4 *__restrict__ t4, int64_t nt,
                                                               const double *__restrict__ x, int nf,
                                                               double *__restrict__ cx) {
    using E = Ex<T>;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nt) return;
    const int4 ti = __ldg(t4 + e);
    const Vec3<T> p0 = load_vertex<T>(v4, ti.x), p1 = load_vertex<T>(v4, ti.y), p2 = load_vertex<T>(v4, ti.z);
    const Vec3<T> e2 = vsub(p1, p0), e0 = vsub(p2, p1), e1 = vsub(p0, p2);
    Vec3<T> nrm = vcross(e2, vneg(e1));
    T ln = E::sqrt(vdot(nrm, nrm));
    if (ln < E::eps()) ln = (T)1;
    const T lni = E::div((T)1, ln);
    nrm = {E::mul(nrm.x, lni), E::mul(nrm.y, lni), E::mul(nrm.z, lni)};
    const Vec3<double> w0 = vwiden(vcross(e0, nrm)), w1 = vwiden(vcross(e1, nrm)), w2 = vwiden(vcross(e2, nrm));
    for (int k = 0; k < nf; k++) {
        const double *xv = x + (e * nf + k) * 3;
        cx[(e * 3 + 0) * nf + k] = dot_wide(w0, xv);
        cx[(e * 3 + 1) * nf + k] = dot_wide(w1, xv);
        cx[(e * 3 + 2) * nf + k] = dot_wide(w2, xv);
    }
}

// g <- g / |g| per 3-vector with NumPy's nan_to_num (diffgeo.py:150-155): (gx^2 + gy^2) + gz^2, one
// IEEE division per component, nan -> 0, +-inf -> +-DBL_MAX
__global__ void normalize_field_kernel(int64_t count, double *__restrict__ g) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double *p = g + 3 * i;
    const double x = p[0], y = p[1], z = p[2];
    const double nrm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
    const double big = 1.7976931348623157e308;
    double o[3] = {__ddiv_rn(x, nrm), __ddiv_rn(y, nrm), __ddiv_rn(z, nrm)};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (o[k] != o[k]) o[k] = 0.0;
        else if (o[k] > big) o[k] = big;
        else if (o[k] < -big) o[k] = -big;
        p[k] = o[k];
    }
}

template <class T>
__global__ void __launch_bounds__(256) tet_div_corner_kernel(const typename Ex<T>::V4 *__restrict__ v4,
                                                             const int4 *__restrict__ t4, int64_t nt,
                                                             const double *__restrict__ x, int nf,
                                                             double *__restrict__ cx) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nt) return;
    const int4 ti = __ldg(t4 + e);
    const Vec3<T> p0 = load_vertex<T>(v4, ti.x), p1 = load_vertex<T>(v4, ti.y);
    const Vec3<T> p2 = load_vertex<T>(v4, ti.z), p3 = load_vertex<T>(v4, ti.w);
    const Vec3<T> e0 = vsub(p1, p0), e1 = vsub(p2, p1), e2 = vsub(p2, p0), e3 = vsub(p3, p0), e4 = vsub(p3, p1);
    const Vec3<double> n0 = vwiden(vcross(e1, e4)), n1 = vwiden(vcross(e3, e2));
    const Vec3<double> n2 = vwiden(vcross(e0, e3)), n3 = vwiden(vcross(e2, e0));
    for (int k = 0; k < nf; k++) {
        const double *xv = x + (e * nf + k) * 3;
        cx[(e * 4 + 0) * nf + k] = dot_wide(n0, xv);
        cx[(e * 4 + 1) * nf + k] = dot_wide(n1, xv);
        cx[(e * 4 + 2) * nf + k] = dot_wide(n2, xv);
        cx[(e * 4 + 3) * nf + k] = dot_wide(n3, xv);
    }
}

// d[v,k] = scale * sum over incident (element, corner), ascending, of cx
__global__ void vertex_gather_kernel(int64_t nv, int kverts, const int32_t *__restrict__ inc_ptr,
                                     const int32_t *__restrict__ inc, const double *__restrict__ cx, int nf,
                                     double scale, double *__restrict__ d) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nv * nf) return;
    const int64_t v = t / nf;
    const int k = (int)(t - v * nf);
    double s = 0.0;
    bool first = true;
    for (int p = inc_ptr[v]; p < inc_ptr[v + 1]; p++) {
        const int code = inc[p];
        const int64_t slot = (int64_t)(code >> 2) * kverts + (code & 3);
        const double val = cx[slot * nf + k];
        s = first ? val : __dadd_rn(s, val);
        first = false;
    }
    d[t] = __dmul_rn(scale, s);
}

// device-level forms (inputs / outputs resident): df (nv, nf) -> dg (nt, nf, 3); dx (nt, nf, 3) -> dd (nv, nf)
static void gradient_dev(lb_ctx *c, lb_mesh *mesh, const double *df, int nf, double *dg) {
    const int64_t nt = mesh->nt;
    const int grid = cdiv(nt, 256);
    if (mesh->k == 3) {
        if (mesh->v_dtype == LB_F32)
            LB_LAUNCH(c, tria_gradient_kernel<float>, grid, 256, 0, mesh->v4f.p, mesh->t4.p, nt, df, nf, dg);
        else
            LB_LAUNCH(c, tria_gradient_kernel<double>, grid, 256, 0, mesh->v4s->p, mesh->t4.p, nt, df, nf, dg);
    } else {
        if (mesh->v_dtype == LB_F32)
            LB_LAUNCH(c, tet_gradient_kernel<float>, grid, 256, 0, mesh->v4f.p, mesh->t4.p, nt, df, nf, dg);
        else
            LB_LAUNCH(c, tet_gradient_kernel<double>, grid, 256, 0, mesh->v4s->p, mesh->t4.p, nt, df, nf, dg);
    }
}

static void divergence_dev(lb_ctx *c, lb_mesh *mesh, const double *dx, int nf, double *dd, bool flux) {
    const int64_t nv = mesh->nv, nt = mesh->nt;
    const int k = mesh->k;
    ensure_incidence(mesh);
    DBuf<double> cx(c, (size_t)nt * k * nf);
    const int grid = cdiv(nt, 256);
    if (k == 3 && flux) {
        if (mesh->v_dtype == LB_F32)
            LB_LAUNCH(c, tria_div2_corner_kernel<float>, grid, 256, 0, mesh->v4f.p, mesh->t4.p, nt, dx, nf, cx.p);
        else
            LB_LAUNCH(c, tria_div2_corner_kernel<double>, grid, 256, 0, mesh->v4s->p, mesh->t4.p, nt, dx, nf, cx.p);
    } else if (k == 3) {
        if (mesh->v_dtype == LB_F32)
            LB_LAUNCH(c, tria_div_corner_kernel<float>, grid, 256, 0, mesh->v4f.p, mesh->t4.p, nt, dx, nf, cx.p);
        else
            LB_LAUNCH(c, tria_div_corner_kernel<double>, grid, 256, 0, mesh->v4s->p, mesh->t4.p, nt, dx, nf, cx.p);
    } else {
        if (mesh->v_dtype == LB_F32)
            LB_LAUNCH(c, tet_div_corner_kernel<float>, grid, 256, 0, mesh->v4f.p, mesh->t4.p, nt, dx, nf, cx.p);
        else
            LB_LAUNCH(c, tet_div_corner_kernel<double>, grid, 256, 0, mesh->v4s->p, mesh->t4.p, nt, dx, nf, cx.p);
    }
    // 0.5 * sum (tria, diffgeo.py:365, :385, :452, :466);  -(1/6) * sum (tet, diffgeo.py:986, :1004)
    const double scale = k == 3 ? 0.5 : -(1.0 / 6.0);
    LB_LAUNCH(c, vertex_gather_kernel, cdiv(nv * nf, 256), 256, 0, nv, k, mesh->inc_ptr.p, mesh->inc.p, cx.p, nf, scale,
              dd);
}

}  // namespace lb

using namespace lb;

extern "C" {

int lb_gradient(lb_ctx *c, lb_mesh *mesh, const double *f, int64_t nf, double *g) {
    LB_API_BEGIN
    LB_REQUIRE(c && mesh && f && g, "lb_gradient: NULL argument");
    LB_REQUIRE(nf >= 1 && nf <= 4096, "lb_gradient: bad number of functions");
    DeviceGuard guard(c->device);
    const int64_t nv = mesh->nv, nt = mesh->nt;
    DBuf<double> df(c, (size_t)nv * nf), dg(c, (size_t)nt * nf * 3);
    h2d(c, df.p, f, (size_t)nv * nf * sizeof(double));
    gradient_dev(c, mesh, df.p, (int)nf, dg.p);
    d2h_large(c, g, dg.p, (size_t)nt * nf * 3 * sizeof(double));
    sync(c);
    LB_API_END
}

static int divergence_api(lb_ctx *c, lb_mesh *mesh, const double *x, int64_t nf, double *d, bool flux) {
    LB_API_BEGIN
    LB_REQUIRE(c && mesh && x && d, "lb_divergence: NULL argument");
    LB_REQUIRE(nf >= 1 && nf <= 4096, "lb_divergence: bad number of functions");
    LB_REQUIRE(!flux || mesh->k == 3, "lb_divergence2 is defined for triangle meshes");
    DeviceGuard guard(c->device);
    const int64_t nv = mesh->nv, nt = mesh->nt;
    DBuf<double> dx(c, (size_t)nt * nf * 3), dd(c, (size_t)nv * nf);
    h2d(c, dx.p, x, (size_t)nt * nf * 3 * sizeof(double));
    divergence_dev(c, mesh, dx.p, (int)nf, dd.p, flux);
    d2h(c, d, dd.p, (size_t)nv * nf * sizeof(double));
    sync(c);
    LB_API_END
}

int lb_divergence(lb_ctx *c, lb_mesh *mesh, const double *x, int64_t nf, double *d) {
    return divergence_api(c, mesh, x, nf, d, false);
}

int lb_divergence2(lb_ctx *c, lb_mesh *mesh, const double *x, int64_t nf, double *d) {
    return divergence_api(c, mesh, x, nf, d, true);
}

// d (nv, nf) = div( grad f / |grad f| ): the right-hand side of the heat method's Poisson problem
// (compute_geodesic_f, lapy/diffgeo.py:144-156) without a host round trip of the (nt, nf, 3) field
int lb_unit_gradient_divergence(lb_ctx *c, lb_mesh *mesh, const double *f, int64_t nf, double *d) {
    LB_API_BEGIN
    LB_REQUIRE(c && mesh && f && d, "lb_unit_gradient_divergence: NULL argument");
    LB_REQUIRE(nf >= 1 && nf <= 4096, "lb_unit_gradient_divergence: bad number of functions");
    DeviceGuard guard(c->device);
    const int64_t nv = mesh->nv, nt = mesh->nt;
    DBuf<double> df(c, (size_t)nv * nf), dg(c, (size_t)nt * nf * 3), dd(c, (size_t)nv * nf);
    h2d(c, df.p, f, (size_t)nv * nf * sizeof(double));
    gradient_dev(c, mesh, df.p, (int)nf, dg.p);
    LB_LAUNCH(c, normalize_field_kernel, cdiv(nt * nf, 256), 256, 0, nt * nf, dg.p);
    divergence_dev(c, mesh, dg.p, (int)nf, dd.p, false);
    d2h(c, d, dd.p, (size_t)nv * nf * sizeof(double));
    sync(c);
    LB_API_END
}

}  // extern "C"
