// Plan-compiler microbenchmark (host only): single-thread pass profile and multi-thread throughput of tb::compile_plan
// on real branch networks.  Built and driven by scripts/plan_bench/run.py; not part of libtbcuda.so.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>

#include "plan.hpp"

#ifdef TB_PLAN_PROFILE
void tb_plan_profile_dump(int reps);
#endif

// best-of-reps time per plan, one thread, plans compiled into a recycled plan object (as the engine's workers do)
extern "C" double pb_single(const tb_network* nets, int n, int reps) {
    double best = 1e30;
    tb::Plan P;
    for (int r = 0; r < reps; ++r) {
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < n; ++i) {
            std::string err;
            P.recycle();
            if (int rc = tb::compile_plan(nets[i], tb::TB_PLAN_TEMPORARY, P, err)) {
                printf("error %d: %s\n", rc, err.c_str());
                return -1;
            }
        }
        auto t1 = std::chrono::steady_clock::now();
        best = std::min(best, std::chrono::duration<double, std::micro>(t1 - t0).count() / double(n));
    }
    return best;
}

// wall time per plan with `threads` workers; every plan's descriptors are kept until the end of the pass, as in a call
extern "C" double pb_threads(const tb_network* nets, int n, int threads, int reps) {
    double best = 1e30;
    std::vector<tb::Plan*> out(n, nullptr);
    for (int r = 0; r < reps; ++r) {
        std::atomic<int> next{0};
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < threads; ++t)
            th.emplace_back([&] {
                tb::Plan scratch;
                for (;;) {
                    int i = next.fetch_add(1);
                    if (i >= n) break;
                    std::string err;
                    scratch.recycle();
                    tb::compile_plan(nets[i], tb::TB_PLAN_TEMPORARY, scratch, err);
                    if (!out[i]) out[i] = new tb::Plan();
                    scratch.copy_descriptors_to(*out[i]);
                }
            });
        for (auto& t : th) t.join();
        auto t1 = std::chrono::steady_clock::now();
        best = std::min(best, std::chrono::duration<double, std::micro>(t1 - t0).count() / double(n));
    }
    for (auto* p : out) delete p;
    return best;
}

extern "C" void pb_profile_dump(int compilations) {
#ifdef TB_PLAN_PROFILE
    tb_plan_profile_dump(compilations);
#else
    (void)compilations;
#endif
}
