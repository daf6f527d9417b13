// C++ mirror of the reference's own tests, through rust_lbfgs_b200/cxx/lbfgsb200.hpp -> the C ABI.
//   tests/simple.rs:17-55   Rosenbrock N=100, then OWL-QN from the converged point      (P2, P3)
//   tests/simple.rs:57-83   Booth function, written as a HOST closure like the reference   (P4)
//   src/lib.rs:9-53         doc-test: with_max_iterations(5) returns Ok                    (P6)
// `test_builder --no-gpu` runs only the checks that need no device (builder asserts, defaults).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../rust_lbfgs_b200/cxx/lbfgsb200.hpp"

using namespace lbfgsb200;

static int failures = 0;
#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); ++failures; } \
    } while (0)

template <class F>
static bool panics(F f) {
    try { f(); } catch (const std::invalid_argument &) { return true; } catch (const std::logic_error &) { return true; }
    return false;
}

static void builder_checks() {
    const lbfgsb200_param_t p = lbfgs().param();                        // src/lbfgs.rs:156-177
    CHECK(p.m == 6 && p.epsilon == 1e-5 && p.max_iterations == 0 && p.ls_max_linesearch == 20);
    CHECK(p.ls_algorithm == LBFGSB200_LS_MORETHUENTE && p.ls_ftol == 1e-4 && p.ls_gtol == 0.9);
    Lbfgs g = lbfgs();
    g.with_gradient_only();                                             // :283-289
    CHECK(g.param().ls_gradient_only == 1 && g.param().damping == 1 && g.param().ls_algorithm == LBFGSB200_LS_BACKTRACKING_STRONG_WOLFE);
    CHECK(panics([] { lbfgs().with_epsilon(-1.0); }));
    CHECK(panics([] { lbfgs().with_max_step_size(-0.0); }));
    CHECK(panics([] { lbfgs().with_linesearch_gtol(1.5); }));
    CHECK(panics([] { lbfgs().with_linesearch_ftol(-1.0); }));
    CHECK(panics([] { lbfgs().with_orthantwise(-1.0, 0); }));
    CHECK(panics([] { lbfgs().with_linesearch_algorithm("Newton"); }));   // unimplemented!()
    CHECK(!panics([] { lbfgs().with_linesearch_algorithm("Backtracking"); }));
}

int main(int argc, char **argv) {
    builder_checks();
    if (argc > 1 && !std::strcmp(argv[1], "--no-gpu")) {
        if (lbfgsb200_device_count() == 0) {   // no fallback: creating anything must fail loudly
            bool threw = false;
            try { Objective::rosenbrock(); } catch (const Error &e) { threw = e.status == LBFGSB200_ERR_CUDA; }
            CHECK(threw);
        }
        std::printf(failures ? "FAILED\n" : "ok (no-gpu)\n");
        return failures != 0;
    }

    // ---- P2: tests/simple.rs:17-40 --------------------------------------------------------------
    const int N = 100;
    std::vector<double> x(N);
    for (int i = 0; i < N; i += 2) { x[i] = -1.2; x[i + 1] = 1.0; }
    Objective rosen = Objective::rosenbrock();
    int ncb = 0;
    Report r = lbfgs().minimize(x, rosen, [&](const Progress &p) { ++ncb; return p.niter < 0; });
    CHECK(std::fabs(r.fx) <= 1e-4);
    for (double v : x) CHECK(std::fabs(v - 1.0) <= 1e-4);
    CHECK(r.status == LBFGSB200_OK_CONVERGED && ncb == 35 && r.neval == 40);   // the oracle's counts
    std::printf("P2 rosenbrock: k=%d neval=%lld fx=%.6e\n", ncb, (long long)r.neval, r.fx);

    // ---- P3: tests/simple.rs:43-54, OWL-QN from the converged x --------------------------------------
    Report r2 = lbfgs().with_orthantwise(1.0, 0, 99).minimize(x, rosen);
    CHECK(std::fabs(r2.fx - 43.5025) <= 1e-4);
    CHECK(std::fabs(x[0] - 0.2500) <= 1e-4 && std::fabs(x[1] - 0.0575) <= 1e-4);
    std::printf("P3 owlqn: neval=%lld fx=%.10f x0=%.6f x1=%.6f\n", (long long)r2.neval, r2.fx, x[0], x[1]);

    // ---- P4: tests/simple.rs:57-83, Booth as a host closure ---------------------------------------
    std::vector<double> xb = {-1.2, 1.0};
    Report r3 = lbfgs().minimize(xb, host_evaluate([](const std::vector<double> &v, std::vector<double> &g, double &fx) {
        const double x1 = v[0], x2 = v[1];
        fx = std::pow(x1 + 2.0 * x2 - 7.0, 2) + std::pow(2.0 * x1 + x2 - 5.0, 2);
        g[0] = 10.0 * x1 + 8.0 * x2 - 34.0;
        g[1] = 8.0 * x1 + 10.0 * x2 - 38.0;
        return true;
    }));
    CHECK(std::fabs(xb[0] - 1.0) <= 1e-6 && std::fabs(xb[1] - 3.0) <= 1e-6);
    std::printf("P4 booth: x=(%.9f, %.9f) neval=%lld\n", xb[0], xb[1], (long long)r3.neval);

    // ---- P6: src/lib.rs:38-50 ------------------------------------------------------------------------
    for (int i = 0; i < N; i += 2) { x[i] = -1.2; x[i + 1] = 1.0; }
    Report r4 = lbfgs().with_max_iterations(5).with_orthantwise(1.0, 0, 99).minimize(x, rosen, default_progress());
    CHECK(r4.status == LBFGSB200_OK_MAX_ITERATIONS && r4.niter == 5);

    // ---- the opt-in compact search direction: same problem, same counts, same answer --------------------------
    for (int i = 0; i < N; i += 2) { x[i] = -1.2; x[i + 1] = 1.0; }
    int ncc = 0;
    Report rc = lbfgs().with_direction(LBFGSB200_DIRECTION_COMPACT).minimize(x, rosen, [&](const Progress &) { ++ncc; return false; });
    CHECK(rc.status == LBFGSB200_OK_CONVERGED && ncc == 35 && rc.neval == 40);
    for (double v : x) CHECK(std::fabs(v - 1.0) <= 1e-4);
    CHECK(panics([&] { std::vector<double> xs(4, 0.5); lbfgs().with_m(33).with_direction(LBFGSB200_DIRECTION_COMPACT).minimize(xs, rosen); }));   // m <= 32

    // ---- device-resident x + the iterative API (src/lbfgs.rs:443-566, src/line.rs:9-32) ------------------
    void *xd = nullptr;
    CHECK(lbfgsb200_device_alloc(0, N * sizeof(double), &xd) == 0);
    for (int i = 0; i < N; i += 2) { x[i] = -1.2; x[i + 1] = 1.0; }
    CHECK(lbfgsb200_copy_h2d(xd, x.data(), N * sizeof(double), nullptr) == 0);
    {
        LbfgsState st = lbfgs().build((double *)xd, N, rosen);
        int k = 0;
        while (!st.is_converged()) { Progress p = st.propagate(); ++k; CHECK(p.niter == k); }
        st.finish();
        CHECK(k == 35 && st.report().neval == 40);
    }
    CHECK(lbfgsb200_copy_d2h(x.data(), xd, N * sizeof(double), nullptr) == 0);
    for (double v : x) CHECK(std::fabs(v - 1.0) <= 1e-4);
    // the same through the compact direction (the state knows which one it runs)
    for (int i = 0; i < N; i += 2) { x[i] = -1.2; x[i + 1] = 1.0; }
    CHECK(lbfgsb200_copy_h2d(xd, x.data(), N * sizeof(double), nullptr) == 0);
    {
        LbfgsState st = lbfgs().with_direction(LBFGSB200_DIRECTION_COMPACT).build((double *)xd, N, rosen);
        CHECK(st.direction() == LBFGSB200_DIRECTION_COMPACT);
        int k = 0;
        while (!st.is_converged()) { st.propagate(); ++k; }
        st.finish();
        CHECK(k == 35 && st.report().neval == 40);
    }
    CHECK(lbfgsb200_copy_d2h(x.data(), xd, N * sizeof(double), nullptr) == 0);
    for (double v : x) CHECK(std::fabs(v - 1.0) <= 1e-4);
    // cancel from the progress callback (src/lbfgs.rs:412-416)
    CHECK(lbfgsb200_copy_h2d(xd, std::vector<double>(N, 0.5).data(), N * sizeof(double), nullptr) == 0);
    Report r5 = lbfgs().minimize((double *)xd, N, rosen, [](const Progress &p) { return p.niter == 3; });
    CHECK(r5.status == LBFGSB200_OK_CANCELLED && r5.niter == 3);
    // an Err from evaluate at the initial point propagates (src/lbfgs.rs:454)
    bool threw = false;
    try {
        lbfgs().minimize((double *)xd, N, DeviceEvaluate([](const double *, double *, int64_t, void *, double *) { return 1; }));
    } catch (const Error &e) { threw = e.status == LBFGSB200_ERR_EVALUATE; }
    CHECK(threw);
    lbfgsb200_device_free(xd);

    std::printf(failures ? "FAILED (%d)\n" : "ok\n", failures);
    return failures != 0;
}
