// fractal_compression.hpp -- C++ mirror of the reference's host-side codec facade on top of
// the C ABI (include/fic_b200.h).  The reference is Java (src/bvk_ss19/FractalCompression.java
// = FC, RasterImage.java = RI); no JDK exists in the build image, so the part that stays on the
// host (image container, grey/RGB dispatch, stream writing, decode entry) is mirrored here with
// the reference's names.  Header only; link with -lfic_b200.
#pragma once

#include <cstdint>
#include <istream>
#include <iterator>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fic_b200.h"

namespace bvk_ss19 {

// RI:18-72
struct RasterImage {
    std::vector<int32_t> argb;  // 0xAARRGGBB, scanline order
    int width = 0, height = 0;
    RasterImage() = default;
    RasterImage(int w, int h) : argb((size_t)w * h, (int32_t)0xffa0a0a0), width(w), height(h) {}  // RI:19, RI:31
};

class FractalCompression {
public:
    static inline int blockgroesse = 8;   // FC:14
    static inline int widthKernel = 2;    // FC:15
    static inline float avgError = 0.0f;  // FC:20 (never reset between decodes)
    static inline int device = 0;
    // Extension, not in the reference: try every candidate domain under the 8 isometries of the square
    // (grey images only; codes gain a 4th int, the stream header carries 2 instead of 0).
    static inline bool isometries = false;

    static float getAvgError() { return avgError; }  // FC:22-24

    // FC:32-45
    static bool isGreyScale(const RasterImage &in)
    {
        for (int32_t p : in.argb) {
            int r = (p >> 16) & 0xff, g = (p >> 8) & 0xff, b = p & 0xff;
            if (r != g || g != b || b != r) return false;
        }
        return true;
    }

    // FC:54-59 -> FC:109-162 / FC:171-219.  Writes the .run stream to `out`, returns the collage.
    static RasterImage encode(const RasterImage &in, std::ostream &out)
    {
        const int rgb = !isGreyScale(in) ? FIC_MODE_RGB : (isometries ? FIC_MODE_GREY_ISO : FIC_MODE_GREY);
        const int S = rgb == FIC_MODE_RGB ? 5 : (rgb == FIC_MODE_GREY_ISO ? 4 : 3);
        int64_t nr = 0;
        check(fic_geometry(in.width, in.height, blockgroesse, widthKernel, &nr, nullptr), "fic_geometry");
        imageInfo.assign((size_t)nr * S, 0.0f);  // FC:124 / FC:185
        std::vector<int32_t> q((size_t)nr * S);
        auto entry = rgb == FIC_MODE_RGB ? fic_encode_rgb : (rgb == FIC_MODE_GREY_ISO ? fic_encode_grey_iso : fic_encode_grey);
        check(entry(handle(), in.argb.data(), in.width, in.height, blockgroesse, widthKernel, 0, nr, imageInfo.data(), q.data()),
              "fic_encode");
        // FC:230-261 writeData
        std::vector<uint8_t> bytes(fic_stream_size(rgb, in.width, in.height, blockgroesse));
        check(fic_stream_write(rgb, in.width, in.height, blockgroesse, widthKernel, q.data(), bytes.data(), bytes.size()),
              "fic_stream_write");
        out.write((const char *)bytes.data(), (std::streamsize)bytes.size());
        out.flush();
        // FC:161 / FC:218 getBestGeneratedCollage[RGB] (rewrites imageInfo[.][0] in place, FC:273)
        RasterImage collage(in.width, in.height);
        check(fic_collage(handle(), rgb, in.argb.data(), in.width, in.height, blockgroesse, widthKernel,
                          imageInfo.data(), collage.argb.data()),
              "fic_collage");
        return collage;
    }

    // FC:547-553 -> FC:356-421 / FC:430-508
    static RasterImage decode(std::istream &in)
    {
        std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
        int rgb, W, H, B, wk;
        size_t off;
        check(fic_stream_read_header(bytes.data(), bytes.size(), &rgb, &W, &H, &B, &wk, &off), "fic_stream_read_header");
        std::vector<int32_t> q((bytes.size() - off) / 4);
        check(fic_stream_read_codes(bytes.data(), bytes.size(), q.data()), "fic_stream_read_codes");
        RasterImage img(W, H);
        int iters = 0;
        check(fic_decode(handle(), rgb, W, H, B, wk, q.data(), 50, img.argb.data(), &avgError, &iters), "fic_decode");
        lastIterations = iters;
        return img;
    }

    static inline std::vector<float> imageInfo;  // FC:17-18, flattened [NR][3|5]
    static inline int lastIterations = 0;

private:
    static fic_handle *handle()
    {
        static fic_handle *h = nullptr;
        if (!h && fic_create(device, &h)) throw std::runtime_error(std::string("fic_create: ") + fic_last_error(nullptr));
        return h;
    }
    static void check(int rc, const char *what)
    {
        // the reference signals every failure as `throws Exception` (FC:54, FC:230, FC:547)
        if (rc) throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + fic_last_error(handle_or_null()));
    }
    static fic_handle *handle_or_null()
    {
        try { return handle(); } catch (...) { return nullptr; }
    }
};

}  // namespace bvk_ss19
