// efficient_host.cpp — host control of the table-based renderer (reference
// src/systems.rs:333-527).  The photon integrations run on the GPU in batches (one launch per
// refinement pass of the sampler); what stays on the host is control flow and a few hundred
// closed-form evaluations per frame, done with the platform libm in the reference's operation
// order (compiled -ffp-contract=off) so the table equals the reference's:
//   compute_escape_angle                    src/systems.rs:203-261
//   escaped_photon_to_world_direction       src/systems.rs:144-187
//   doubly_sample_function & helpers        src/sampling.rs:21-245
//   interp::interp_slice (interp 1.0.3, Cargo.lock:474-475; published behaviour: per-segment
//     slope dy/dx (0 when dx == 0), intercept y - x*m, last index strictly below xp clamped to
//     len-2, linear extrapolation beyond both ends)
// nalgebra 0.33.0 (Cargo.lock:623-624) Rotation3::rotation_between / from_axis_angle restated
// from their published behaviour.
#include "efficient.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <unordered_map>
#include "launch_host.h"

namespace curvis {

namespace {

constexpr double kPi = 3.14159265358979323846264338327950288;

inline double norm3(const double v[3]) { return std::sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); }
inline double dot3(const double a[3], const double b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
inline void cross3(const double a[3], const double b[3], double o[3]) {
    const double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
inline void mat_vec(const double m[9], const double v[3], double o[3]) {
    for (int i = 0; i < 3; ++i) o[i] = (m[3 * i] * v[0] + m[3 * i + 1] * v[1]) + m[3 * i + 2] * v[2];
}
inline void identity(double m[9]) { m[0] = 1; m[1] = 0; m[2] = 0; m[3] = 0; m[4] = 1; m[5] = 0; m[6] = 0; m[7] = 0; m[8] = 1; }

// f64::rem_euclid
inline double rem_euclid(double x, double rhs) { const double r = std::fmod(x, rhs); return r < 0.0 ? r + std::fabs(rhs) : r; }

// Rotation3::from_axis_angle
void from_axis_angle(const double u[3], double angle, double m[9]) {
    if (!(angle != 0.0)) { identity(m); return; }
    const double ux = u[0], uy = u[1], uz = u[2];
    const double sqx = ux * ux, sqy = uy * uy, sqz = uz * uz;
    const double sn = std::sin(angle), cs = std::cos(angle), omc = 1.0 - cs;
    m[0] = sqx + (1.0 - sqx) * cs;      m[1] = ux * uy * omc - uz * sn;  m[2] = ux * uz * omc + uy * sn;
    m[3] = ux * uy * omc + uz * sn;     m[4] = sqy + (1.0 - sqy) * cs;   m[5] = uy * uz * omc - ux * sn;
    m[6] = ux * uz * omc - uy * sn;     m[7] = uy * uz * omc + ux * sn;  m[8] = sqz + (1.0 - sqz) * cs;
}

double shape_r_squared(const curvis_metric& g, double l) {
    switch (g.kind) {
    case CURVIS_METRIC_ELLIS: return g.rho * g.rho + l * l;                       // metrics.rs:419
    case CURVIS_METRIC_INTERSTELLAR: { const double r = host_shape_r(g, l); return r * r; }   // :474
    default: return l * l;                                                        // :503
    }
}

struct Sample { double a, e, s; };

inline bool finite3(const Sample& p) { return std::isfinite(p.a) && std::isfinite(p.e) && std::isfinite(p.s); }
void clean(std::vector<Sample>& v) {                                              // sampling.rs:21-32
    size_t k = 0;
    for (size_t i = 0; i < v.size(); ++i) if (finite3(v[i])) v[k++] = v[i];
    v.resize(k);
}

// The tail of compute_escape_angle (systems.rs:236-259) on a photon the device integrated.
// Returns false where the reference would panic (rotation_from_two_vectors on parallel vectors).
bool escape_angle_from_record(const curvis_metric& g, const curvis_ray_record& rec, double& angle, double& sign) {
    if (rec.side == 0) { angle = NAN; sign = NAN; return true; }                  // EscapeAngle::NotEscaped -> (NaN, NaN), :484
    // relativistic_vector_to_direction (metrics.rs:339-349), covariant momentum
    const double s = std::sin(rec.theta);
    const double r2 = shape_r_squared(g, rec.l), r = host_shape_r(g, rec.l);
    const double tangent[3] = {(rec.p_l * (1.0 / 1.0)) * 1.0, (rec.p_theta * (1.0 / r2)) * r, (rec.p_phi * (1.0 / (r2 * (s * s)))) * r};
    double world_position[3], rot[9], wd[3];
    host_vector3_from_theta_phi(rec.theta, rec.phi, world_position);              // systems.rs:176
    const double ex[3] = {1.0, 0.0, 0.0};
    if (!host_rotation_from_two_vectors(ex, world_position, rot)) return false;   // :178-181
    mat_vec(rot, tangent, wd);                                                    // :183
    const double n = norm3(wd);                                                   // :246
    wd[0] = wd[0] / n; wd[1] = wd[1] / n; wd[2] = wd[2] / n;
    const double vx = (wd[0] * 1.0 + wd[1] * 0.0) + wd[2] * 0.0;
    const double vy = (wd[0] * 0.0 + wd[1] * 1.0) + wd[2] * 0.0;
    angle = (vy >= 0.0) ? std::acos(vx) : 2.0 * kPi - std::acos(vx);              // :251
    sign = rec.side > 0 ? 1.0 : -1.0;                                             // :481-483
    return true;
}

}  // namespace

void host_vector3_from_theta_phi(double theta, double phi, double out[3]) {       // algebra.rs:106-126
    if (theta < 0.0) { theta = std::fabs(theta); phi = phi + kPi; }
    phi = rem_euclid(phi, 2.0 * kPi);
    out[0] = std::sin(theta) * std::cos(phi);
    out[1] = std::sin(theta) * std::sin(phi);
    out[2] = std::cos(theta);
}

bool host_rotation_from_two_vectors(const double v1[3], const double v2[3], double m[9]) {   // algebra.rs:92-101
    double c[3];
    cross3(v1, v2, c);
    if (norm3(c) == 0.0) return false;                                            // :95-97 panic
    const double n1 = norm3(v1), n2 = norm3(v2);
    if (n1 > 0.0 && n2 > 0.0) {                                                   // Rotation3::rotation_between
        const double na[3] = {v1[0] / n1, v1[1] / n1, v1[2] / n1}, nb[3] = {v2[0] / n2, v2[1] / n2, v2[2] / n2};
        cross3(na, nb, c);
        const double cn = norm3(c);
        if (cn > 2.220446049250313e-16) {
            const double axis[3] = {c[0] / cn, c[1] / cn, c[2] / cn};
            from_axis_angle(axis, std::acos(dot3(na, nb)) * 1.0, m);
            return true;
        }
        if (dot3(na, nb) < 0.0) return false;                                     // None.unwrap()
    }
    identity(m);
    return true;
}

int build_escape_table(const curvis_metric& metric, double l_camera, uint32_t alphas_num, uint32_t max_iterations_sampling,
                       double threshold_1, double threshold_2, const BatchIntegrate& integrate, EscapeTable& table, std::string& err) {
    (void)l_camera;
    table = EscapeTable();
    std::vector<double> dirs;
    std::vector<curvis_ray_record> recs;
    bool panicked = false;
    // expensive_function over a batch of alphas (the closure at systems.rs:470-485).  Results are
    // cached by the exact bit pattern of alpha, and every launch also integrates, speculatively,
    // the midpoints the NEXT refinement levels could ask for inside the segments being refined
    // (kLookahead levels): the sampler's passes are sequential and latency-bound (a few hundred
    // photons cannot fill the GPU), so three levels per launch cut the launch count ~3x without
    // changing a single table value.  Only consumed evaluations are counted in the statistics.
    struct Cached { double e, s; uint32_t steps; };
    std::unordered_map<uint64_t, Cached> cache;
    auto key = [](double a) { uint64_t k; std::memcpy(&k, &a, sizeof k); return k; };
    constexpr int kLookahead = 3;
    std::function<void(double, double, int, std::vector<double>&)> speculate = [&](double lo, double hi, int depth, std::vector<double>& out) {
        if (depth == 0) return;
        const double mid = (lo + hi) / 2.0;                                       // the expression of sampling.rs:176-177
        if (!(mid > lo && mid < hi)) return;
        if (!cache.count(key(mid))) out.push_back(mid);
        speculate(lo, mid, depth - 1, out);
        speculate(mid, hi, depth - 1, out);
    };
    auto integrate_into_cache = [&](std::vector<double>& alphas) -> int {
        std::sort(alphas.begin(), alphas.end());
        alphas.erase(std::unique(alphas.begin(), alphas.end()), alphas.end());
        if (alphas.empty()) return CURVIS_OK;
        dirs.resize(alphas.size() * 3);
        recs.resize(alphas.size());
        for (size_t i = 0; i < alphas.size(); ++i) {                              // systems.rs:221
            dirs[3 * i] = std::cos(alphas[i]); dirs[3 * i + 1] = 0.0; dirs[3 * i + 2] = std::sin(alphas[i]);
        }
        const int rc = integrate(dirs.data(), alphas.size(), recs.data());
        if (rc != CURVIS_OK) return rc;
        for (size_t i = 0; i < alphas.size(); ++i) {
            Cached c;
            if (!escape_angle_from_record(metric, recs[i], c.e, c.s)) { panicked = true; c.e = NAN; c.s = NAN; }
            c.steps = recs[i].steps;
            cache[key(alphas[i])] = c;
        }
        return CURVIS_OK;
    };
    // `segments`: for each pair of new alphas, the (b1, b2, b3) triple they subdivide (empty for the initial range)
    auto evaluate = [&](const std::vector<double>& alphas, const std::vector<double>& segments, std::vector<Sample>& out) -> int {
        std::vector<double> todo;
        for (double a : alphas) if (!cache.count(key(a))) todo.push_back(a);
        if (!todo.empty()) {
            for (size_t t = 0; t + 3 <= segments.size(); t += 3) {
                // children of the four sub-segments (b1,m1) (m1,b2) (b2,m2) (m2,b3)
                const double b1 = segments[t], b2 = segments[t + 1], b3 = segments[t + 2];
                const double m1 = (b1 + b2) / 2.0, m2 = (b2 + b3) / 2.0;
                speculate(b1, m1, kLookahead, todo); speculate(m1, b2, kLookahead, todo);
                speculate(b2, m2, kLookahead, todo); speculate(m2, b3, kLookahead, todo);
            }
            const int rc = integrate_into_cache(todo);
            if (rc != CURVIS_OK) return rc;
            table.passes += 1;                                                    // device launches
        }
        out.resize(alphas.size());
        for (size_t i = 0; i < alphas.size(); ++i) {
            const Cached& c = cache[key(alphas[i])];
            out[i].a = alphas[i]; out[i].e = c.e; out[i].s = c.s;
            table.evaluations += 1;
            table.steps += c.steps;
        }
        return CURVIS_OK;
    };

    const double a_min = -0.1 * kPi, a_max = 1.1 * kPi;                           // systems.rs:437-438
    std::vector<double> xs(alphas_num);
    const double step = (a_max - a_min) / ((double)((size_t)alphas_num - 1));     // sampling.rs:135
    for (uint32_t i = 0; i < alphas_num; ++i) xs[i] = a_min + (double)i * step;
    std::vector<Sample> cur, fresh;
    int rc = evaluate(xs, {}, cur);
    if (rc != CURVIS_OK) { err = "device integration failed while sampling"; return rc; }
    clean(cur);

    uint32_t iteration = 0;
    while (iteration < max_iterations_sampling) {                                 // sampling.rs:90
        const size_t previous = cur.size();
        clean(cur);
        if (cur.size() < 3) {                                                     // :158-160 panic
            err = "bipoints list has length < 3. Cannot proceed to evaluate denser bipoints.";
            return CURVIS_ERR_INVALID_ARGUMENT;
        }
        // pass 1: which triples refine (decisions read existing points only), collect the new alphas
        std::vector<size_t> starts;      // i of each visited triple
        std::vector<char> refine;
        std::vector<double> new_alphas, segments;
        for (size_t i = 0; i < cur.size() - 2;) {
            const Sample &b1 = cur[i], &b2 = cur[(i + 1) % cur.size()], &b3 = cur[(i + 2) % cur.size()];
            const double s1 = std::fabs(((b1.a * b2.e + b2.a * b3.e) + b3.a * b1.e) - ((b1.e * b2.a + b2.e * b3.a) + b3.e * b1.a));
            const double s2 = std::fabs(((b1.a * b2.s + b2.a * b3.s) + b3.a * b1.s) - ((b1.s * b2.a + b2.s * b3.a) + b3.s * b1.a));
            starts.push_back(i);
            if (!(s1 > threshold_1 || s2 > threshold_2)) { refine.push_back(0); i += 1; }
            else {
                refine.push_back(1);
                new_alphas.push_back((b1.a + b2.a) / 2.0);                        // :176-177
                new_alphas.push_back((b2.a + b3.a) / 2.0);
                segments.push_back(b1.a); segments.push_back(b2.a); segments.push_back(b3.a);
                i += 2;
            }
        }
        rc = evaluate(new_alphas, segments, fresh);                               // at most one launch per pass
        if (rc != CURVIS_OK) { err = "device integration failed while sampling"; return rc; }
        // pass 2: assemble in the reference's order
        std::vector<Sample> next;
        next.reserve(cur.size() + fresh.size());
        size_t k = 0;
        for (size_t t = 0; t < starts.size(); ++t) {
            const size_t i = starts[t];
            next.push_back(cur[i]);
            if (refine[t]) {
                next.push_back(fresh[k]);
                next.push_back(cur[(i + 1) % cur.size()]);
                next.push_back(fresh[k + 1]);
                k += 2;
            }
        }
        clean(next);
        cur.swap(next);
        table.refinement_passes += 1;
        if (cur.size() < previous) break;                                         // :97-102
        if (cur.size() == previous) break;                                        // :105-107
        iteration += 1;
    }
    if (panicked) { err = "v1 and v2 must not be parallel"; return CURVIS_ERR_PARALLEL_VECTORS; }

    const size_t n = cur.size();
    table.alphas.resize(n); table.escapes.resize(n); table.signs.resize(n);
    for (size_t i = 0; i < n; ++i) { table.alphas[i] = cur[i].a; table.escapes[i] = cur[i].e; table.signs[i] = cur[i].s; }
    // interp_slice segments.  n == 0 -> constant 0, n == 1 -> constant y[0]: one pseudo-segment.
    const size_t segs = n >= 2 ? n - 1 : 1;
    table.m_e.assign(segs, 0.0); table.c_e.assign(segs, 0.0); table.m_s.assign(segs, 0.0); table.c_s.assign(segs, 0.0);
    if (n == 1) { table.c_e[0] = cur[0].e; table.c_s[0] = cur[0].s; }
    for (size_t i = 0; i + 1 < n; ++i) {
        const double dx = cur[i + 1].a - cur[i].a;
        const double dye = cur[i + 1].e - cur[i].e, dys = cur[i + 1].s - cur[i].s;
        table.m_e[i] = (dx == 0.0) ? 0.0 : dye / dx;
        table.m_s[i] = (dx == 0.0) ? 0.0 : dys / dx;
        table.c_e[i] = cur[i].e - cur[i].a * table.m_e[i];
        table.c_s[i] = cur[i].s - cur[i].a * table.m_s[i];
    }
    for (size_t i = 0; i + 1 < n; ++i)
        if (!(cur[i].a <= cur[i + 1].a)) { err = "sampled alphas are not sorted"; return CURVIS_ERR_INVALID_ARGUMENT; }
    return CURVIS_OK;
}

}  // namespace curvis
