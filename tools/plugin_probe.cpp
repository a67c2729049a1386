// tools/plugin_probe.cpp — measurement harness, not product: drives the per-handle plugin calls (Evaluator_Multiply then
// Evaluator_Relinearize on opaque handles) from N native threads, the way sunscreen_runtime's rayon workers do
// (sunscreen_runtime/src/run.rs:243,279,415-469), without an interpreter lock between the calls.  The function pointers are
// passed in by the caller (bench.py hands over the addresses from the loaded libb200bfv.so), so this file links nothing.
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

typedef long (*op5_fn)(void *, void *, void *, void *, void *);

extern "C" double plugin_probe_mul_relin(op5_fn multiply, op5_fn relinearize, void *evaluator, void **a, void **b, void **prod,
                                         void **out, void *relin_keys, int pairs, int threads, long *first_error)
{
    std::atomic<long> err{ 0 };
    std::atomic<int> next{ 0 };
    auto work = [&]() {
        for (;;)
        {
            const int i = next.fetch_add(1); // work stealing over the pairs, like a rayon scope
            if (i >= pairs || err.load())
                return;
            long rc = multiply(evaluator, a[i], b[i], prod[i], nullptr);
            if (!rc)
                rc = relinearize(evaluator, prod[i], relin_keys, out[i], nullptr);
            if (rc)
            {
                long zero = 0;
                err.compare_exchange_strong(zero, rc);
            }
        }
    };
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> ts;
    for (int t = 0; t < threads; t++)
        ts.emplace_back(work);
    for (auto &t : ts)
        t.join();
    const auto t1 = std::chrono::steady_clock::now();
    if (first_error)
        *first_error = err.load();
    return std::chrono::duration<double>(t1 - t0).count();
}
