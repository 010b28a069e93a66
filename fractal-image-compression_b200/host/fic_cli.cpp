// fic_cli -- headless stand-in for the reference's GUI action openDecodedImage
// (RLEAppController.java:172-188): encode an image to a .run stream, decode a stream.
//   fic_cli encode in.pgm|in.ppm out.run [blockgroesse] [widthKernel]   (binary P5 / P6)
//   fic_cli decode in.run out.pgm|out.ppm
//   fic_cli roundtrip in.pgm|in.ppm [blockgroesse] [widthKernel]        (prints "MSE <avgError>" like the GUI label)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "fractal_compression.hpp"

using namespace bvk_ss19;

static RasterImage read_pnm(const char *path)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    std::string magic;
    int w, h, maxv;
    f >> magic >> w >> h >> maxv;
    f.get();
    if ((magic != "P5" && magic != "P6") || maxv != 255) throw std::runtime_error("need binary P5/P6 with maxval 255");
    RasterImage img(w, h);
    std::vector<unsigned char> buf((size_t)w * h * (magic == "P6" ? 3 : 1));
    f.read((char *)buf.data(), (std::streamsize)buf.size());
    for (size_t i = 0; i < (size_t)w * h; i++) {
        unsigned r, g, b;
        if (magic == "P6") { r = buf[3 * i]; g = buf[3 * i + 1]; b = buf[3 * i + 2]; }
        else r = g = b = buf[i];
        img.argb[i] = (int32_t)(0xff000000u | (r << 16) | (g << 8) | b);
    }
    return img;
}

static void write_pnm(const char *path, const RasterImage &img)
{
    bool grey = FractalCompression::isGreyScale(img);
    std::ofstream f(path, std::ios::binary);
    f << (grey ? "P5" : "P6") << "\n" << img.width << " " << img.height << "\n255\n";
    for (int32_t p : img.argb) {
        unsigned char rgb[3] = {(unsigned char)((p >> 16) & 0xff), (unsigned char)((p >> 8) & 0xff), (unsigned char)(p & 0xff)};
        f.write((const char *)rgb, grey ? 1 : 3);
    }
}

int main(int argc, char **argv)
{
    // a trailing "iso" selects the isometry extension for encode / roundtrip
    if (argc > 2 && !strcmp(argv[argc - 1], "iso")) {
        FractalCompression::isometries = true;
        argc--;
    }
    try {
        if (argc >= 4 && !strcmp(argv[1], "encode")) {
            if (argc > 4) FractalCompression::blockgroesse = atoi(argv[4]);
            if (argc > 5) FractalCompression::widthKernel = atoi(argv[5]);
            std::ofstream out(argv[3], std::ios::binary);
            FractalCompression::encode(read_pnm(argv[2]), out);
            return 0;
        }
        if (argc >= 4 && !strcmp(argv[1], "decode")) {
            std::ifstream in(argv[2], std::ios::binary);
            write_pnm(argv[3], FractalCompression::decode(in));
            printf("MSE %.9g (%d iterations)\n", FractalCompression::getAvgError(), FractalCompression::lastIterations);
            return 0;
        }
        if (argc >= 3 && !strcmp(argv[1], "roundtrip")) {
            if (argc > 3) FractalCompression::blockgroesse = atoi(argv[3]);
            if (argc > 4) FractalCompression::widthKernel = atoi(argv[4]);
            std::stringstream ss(std::ios::in | std::ios::out | std::ios::binary);
            FractalCompression::encode(read_pnm(argv[2]), ss);
            ss.seekg(0);
            FractalCompression::decode(ss);
            printf("MSE %.9g (%d iterations)\n", FractalCompression::getAvgError(), FractalCompression::lastIterations);
            return 0;
        }
        fprintf(stderr, "usage: fic_cli encode in.pnm out.run [B] [wk] [iso] | decode in.run out.pnm | roundtrip in.pnm [B] [wk] [iso]\n");
        return 2;
    } catch (const std::exception &e) {
        fprintf(stderr, "fic_cli: %s\n", e.what());
        return 1;
    }
}
