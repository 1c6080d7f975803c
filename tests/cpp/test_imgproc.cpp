// test_imgproc.cpp -- drives host/imgproc.hpp + host/imageio.hpp for the cv2 parity tests (tests/test_cli.py).
//   test_imgproc yuv <in.png> <out.f32>                  imread -> convertTo(1/255) -> RGB2YUV, dump float [h][w][3]
//   test_imgproc resize <in.f32> <w> <h> <dw> <dh> <nearest|linear|cubic> <out.f32>
//   test_imgproc rgb8 <in.f32> <w> <h> <out.png>         YUV2RGB -> convertTo(8U,255) -> imwrite
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>

#include "imageio.hpp"
#include "imgproc.hpp"

static w2ximg::Image3f load(const char *path, int w, int h) {
    w2ximg::Image3f im(w, h);
    std::ifstream f(path, std::ios::binary);
    f.read(reinterpret_cast<char *>(im.data.data()), (std::streamsize)(im.data.size() * 4));
    return im;
}
static void dump(const char *path, const w2ximg::Image3f &im) {
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char *>(im.data.data()), (std::streamsize)(im.data.size() * 4));
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    std::string op = argv[1];
    if (op == "yuv" && argc == 4) {
        std::string err;
        w2xio::Image8 in = w2xio::imread(argv[2], &err);
        if (in.empty()) { std::fprintf(stderr, "%s\n", err.c_str()); return 1; }
        w2ximg::Image3f im = w2ximg::from_u8(in.bgr.data(), in.width, in.height);
        w2ximg::rgb2yuv(im);
        dump(argv[3], im);
        std::printf("%d %d\n", in.width, in.height);
        return 0;
    }
    if (op == "resize" && argc == 9) {
        w2ximg::Image3f im = load(argv[2], std::atoi(argv[3]), std::atoi(argv[4]));
        w2ximg::Interp it = !std::strcmp(argv[7], "nearest") ? w2ximg::NEAREST : !std::strcmp(argv[7], "linear") ? w2ximg::LINEAR : w2ximg::CUBIC;
        dump(argv[8], w2ximg::resize(im, std::atoi(argv[5]), std::atoi(argv[6]), it));
        return 0;
    }
    if (op == "rgb8" && argc == 6) {
        w2ximg::Image3f im = load(argv[2], std::atoi(argv[3]), std::atoi(argv[4]));
        w2ximg::yuv2rgb(im);
        std::vector<uint8_t> u8 = w2ximg::to_u8(im);
        return w2xio::imwrite(argv[5], u8.data(), im.width, im.height) ? 0 : 1;
    }
    return 2;
}
