// Static description of the reference network (model/centernet.py:207-261) and of the packed
// weight blob produced by weights.py.  Host-only.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

namespace cf {

struct MBBlock {
    int cin, cout, t, k, s;
    bool residual() const { return cin == cout && s == 1; }  // model/centernet.py:101
    int hid() const { return cin * t; }
};

// settings [t,c,n,s,k] expanded to the 12 MBConv blocks, model/centernet.py:211-234
static const MBBlock kBlocks[12] = {
    {32, 16, 1, 3, 1},                            // layer0
    {16, 24, 6, 3, 2},  {24, 24, 6, 3, 1},        // layer1  -> x1 (stride 4)
    {24, 32, 6, 5, 2},  {32, 32, 6, 5, 1},        // layer2  -> x2 (stride 8)
    {32, 64, 6, 3, 2},  {64, 64, 6, 3, 1},        // layer3
    {64, 96, 6, 5, 1},  {96, 96, 6, 5, 1},        // layer4  -> x4 (stride 16)
    {96, 160, 6, 5, 2}, {160, 160, 6, 5, 1},      // layer5
    {160, 320, 6, 3, 1},                          // layer6
};
static const int kLayerOfBlock[12] = {0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6};
static const int kLastBlockOfLayer[7] = {0, 2, 4, 6, 8, 10, 11};

// ---- weight blob: header | entry table | fp32 payload ----------------------------------
struct BlobHeader {
    char magic[8];  // "CFB200W1"
    uint32_t version;
    uint32_t n_entries;
    uint64_t payload_offset;  // bytes from blob start
    uint64_t payload_floats;
};
struct BlobEntry {
    char name[40];
    uint64_t offset;  // in floats from payload start
    uint64_t count;   // in floats
};

struct Blob {
    const float* payload = nullptr;
    std::map<std::string, std::pair<uint64_t, uint64_t>> idx;
    bool parse(const void* p, size_t bytes, std::string& why) {
        if (bytes < sizeof(BlobHeader)) return why = "blob shorter than its header", false;
        BlobHeader h;
        memcpy(&h, p, sizeof h);
        if (memcmp(h.magic, "CFB200W1", 8) != 0 || h.version != 1) return why = "bad magic/version", false;
        const size_t tab = sizeof(BlobHeader) + (size_t)h.n_entries * sizeof(BlobEntry);
        if (tab > bytes || h.payload_offset < tab || h.payload_offset % 16 != 0 ||
            h.payload_offset + h.payload_floats * 4 > bytes)
            return why = "inconsistent table/payload sizes", false;
        payload = reinterpret_cast<const float*>((const char*)p + h.payload_offset);
        for (uint32_t i = 0; i < h.n_entries; ++i) {
            BlobEntry e;
            memcpy(&e, (const char*)p + sizeof(BlobHeader) + i * sizeof(BlobEntry), sizeof e);
            e.name[39] = 0;
            if (e.offset + e.count > h.payload_floats) return why = std::string("entry out of range: ") + e.name, false;
            idx[e.name] = {e.offset, e.count};
        }
        return true;
    }
    const float* get(const std::string& name, uint64_t count, std::string& why) const {
        auto it = idx.find(name);
        if (it == idx.end()) return why = "missing weight entry " + name, nullptr;
        if (it->second.second != count)
            return why = "entry " + name + " has " + std::to_string(it->second.second) + " floats, expected " +
                         std::to_string(count),
                   nullptr;
        return payload + it->second.first;
    }
};

// ---- cv2.resize INTER_LINEAR tables (OpenCV resize.cpp: the xofs/ialpha and yofs/ibeta loops of cv::resize) -----------
// Same IEEE operations as OpenCV: scale = 1 / (dst / src) in double, f = float((d + 0.5) * scale - 0.5), floor,
// weights rounded half-to-even to 11-bit fixed point.  x clamps (index, fraction) at the borders; y keeps the
// fraction and clips the two row indices.
struct ResizeTables {
    int sh = 0, sw = 0, dh = 0, dw = 0, area2 = 0;
    std::vector<int32_t> tab;  // xofs[dw] | xa0[dw] | xa1[dw] | y0[dh] | y1[dh] | yb0[dh] | yb1[dh]
    void build(int sh_, int sw_, int dh_, int dw_) {
        sh = sh_, sw = sw_, dh = dh_, dw = dw_;
        area2 = (sh == 2 * dh && sw == 2 * dw) ? 1 : 0;
        tab.assign((size_t)3 * dw + (size_t)4 * dh, 0);
        const double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
        for (int dx = 0; dx < dw; ++dx) {
            float fx = (float)((dx + 0.5) * scale_x - 0.5);
            int sx = (int)floorf(fx);
            fx -= sx;
            if (sx < 0) fx = 0.f, sx = 0;
            if (sx >= sw - 1) fx = 0.f, sx = sw - 1;
            tab[dx] = sx;
            tab[dw + dx] = (int32_t)lrintf((1.f - fx) * 2048.f);
            tab[2 * dw + dx] = (int32_t)lrintf(fx * 2048.f);
        }
        int32_t* ty = tab.data() + 3 * dw;
        for (int dy = 0; dy < dh; ++dy) {
            float fy = (float)((dy + 0.5) * scale_y - 0.5);
            const int sy = (int)floorf(fy);
            fy -= sy;
            ty[dy] = sy < 0 ? 0 : (sy > sh - 1 ? sh - 1 : sy);
            ty[dh + dy] = sy + 1 < 0 ? 0 : (sy + 1 > sh - 1 ? sh - 1 : sy + 1);
            ty[2 * dh + dy] = (int32_t)lrintf((1.f - fy) * 2048.f);
            ty[3 * dh + dy] = (int32_t)lrintf(fy * 2048.f);
        }
    }
};

}  // namespace cf
