// bench_host.cpp -- times the API a reference maintainer would actually link (INTEGRATION.md option A/B): the
// w2xc::convertWithModels re-creation in host/w2xc.hpp, planes as w2xc::Plane, progress lines on stdout like the reference
// (src/convertRoutine.cpp:67,133-134), host buffers in, host buffers out.  bench.py runs it for its "e2e_cpp" leg.
//
//   w2x-bench-host <model.json> <width> <height> <steps> <warmup> [gpus]
//
// Prints the reference's progress lines (discard them) and, last, one line "BENCH_JSON {...}".
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "w2xc.hpp"

int main(int argc, char **argv) {
    if (argc < 6) {
        std::fprintf(stderr, "usage: %s model.json width height steps warmup [gpus]\n", argv[0]);
        return 2;
    }
    const std::string model = argv[1];
    const int w = std::atoi(argv[2]), h = std::atoi(argv[3]), steps = std::atoi(argv[4]), warmup = std::atoi(argv[5]);
    if (argc > 6 && std::atoi(argv[6]) > 1) w2xc::gpuRuntime::setNumberOfGpus(std::atoi(argv[6]));
    std::vector<std::unique_ptr<w2xc::Model>> models;
    if (!w2xc::modelUtility::generateModelFromJSON(model, models)) return 1;
    w2xc::Plane in(w, h), out;
    uint64_t s = 0x9E3779B97F4A7C15ull;                       // uniform [0,1) noise (the tolerance-relevant distribution)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            in.at(y, x) = (float)((s >> 40) * (1.0 / 16777216.0));
        }
    for (int i = 0; i < warmup; i++)
        if (!w2xc::convertWithModels(in, out, models)) return 1;
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < steps; i++)
        if (!w2xc::convertWithModels(in, out, models)) return 1;
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / steps;
    double sum = 0.0;
    for (int y = 0; y < h; y += 97) sum += out.at(y, (y * 31) % w);
    std::printf("BENCH_JSON {\"api\": \"w2xc::convertWithModels (host/w2xc.hpp over the C ABI)\", \"width\": %d, \"height\": %d, \"steps\": %d, "
                "\"ms_per_step\": %.4f, \"mpix_per_s\": %.2f, \"checksum\": %.6f}\n",
                w, h, steps, ms, (double)w * h / (ms * 1e-3) / 1e6, sum);
    return 0;
}
