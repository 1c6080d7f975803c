// test_host_api.cpp -- exercises the C++ mirror of the reference interface (host/w2xc.hpp) the way
// the reference's main.cpp / test.cpp use it (src/main.cpp:79-96, src/test.cpp:27,76).
//   test_host_api <model.json> [--gpu <in.f32> <w> <h> <out.f32>]
// Exit code 0 on success; prints CHECK lines.
#include <cstdio>
#include <cstring>
#include <fstream>

#include "w2xc.hpp"

static int fails = 0;
#define CHECK(c) do { if (!(c)) { std::printf("CHECK FAILED line %d: %s\n", __LINE__, #c); fails++; } } while (0)

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    std::vector<std::unique_ptr<w2xc::Model>> models;
    CHECK(!w2xc::modelUtility::generateModelFromJSON("/nonexistent/model.json", models));   // prints "couldn't open"
    CHECK(models.empty());
    CHECK(w2xc::modelUtility::generateModelFromJSON(argv[1], models));
    CHECK(models.size() == 7);
    const int dims[7][2] = {{1, 32}, {32, 32}, {32, 64}, {64, 64}, {64, 128}, {128, 128}, {128, 1}};
    for (size_t i = 0; i < models.size() && i < 7; i++) {
        CHECK(models[i]->getNInputPlanes() == dims[i][0]);
        CHECK(models[i]->getNOutputPlanes() == dims[i][1]);
    }
    auto &mu = w2xc::modelUtility::getInstance();
    CHECK(mu.getNumberOfJobs() == 4);
    CHECK(mu.setNumberOfJobs(2) && mu.getNumberOfJobs() == 2);
    CHECK(!mu.setNumberOfJobs(0));
    int bw = 0, bh = 0;
    mu.getBlockSize(bw, bh);
    CHECK(bw == 512 && bh == 512);
    CHECK(mu.setBlockSizeExp2Square(9));
    if (argc >= 7 && !std::strcmp(argv[2], "--gpu")) {
        const int w = std::atoi(argv[4]), h = std::atoi(argv[5]);
        w2xc::Plane big(w + 10, h + 6);
        w2xc::Plane in = big.roi(3, 2, w, h);                                  // a strided ROI, like the block split path
        std::ifstream f(argv[3], std::ios::binary);
        for (int y = 0; y < h; y++) f.read(reinterpret_cast<char *>(&in.at(y, 0)), (std::streamsize)w * 4);
        w2xc::Plane out;
        CHECK(w2xc::convertWithModels(in, out, models));                       // prints "Iteration #k..."
        CHECK(out.width == w && out.height == h);
        std::ofstream o(argv[6], std::ios::binary);
        for (int y = 0; y < h; y++) o.write(reinterpret_cast<const char *>(&out.at(y, 0)), (std::streamsize)w * 4);
        // Model::filter: wrong plane count -> false + message
        std::vector<w2xc::Plane> ins(3, w2xc::Plane(8, 8)), outs;
        CHECK(!models[1]->filter(ins, outs));
        std::vector<w2xc::Plane> one(1, in.clone());
        CHECK(models[0]->filter(one, outs));
        CHECK(outs.size() == 32 && outs[0].width == w && outs[0].height == h);
    }
    std::printf(fails ? "FAILED %d\n" : "ALL OK\n", fails);
    return fails ? 1 : 0;
}
