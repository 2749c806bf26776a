// Host-side layer table of the Darknet-53 / YOLOv3 conv stack (C++ twin of arch.py).
// Follows src/space/yolov3_detect.py:196-311 (_conv_block, make_yolov3_model) and
// src/space/face_detection.py:341-352, 405-595 (backbone + 3x3x6 head) of the reference.
#pragma once
#include <vector>

namespace fvy {

struct ConvSpec {
    int idx, cin, cout, k, stride;
    bool bn, leaky;
    int src;     // conv idx feeding this conv; -1 network input; -2 concat A (up(84), skip_61); -3 concat B (up(96), skip_36);
                 // -4 a padded bf16 buffer filled by the caller's pack step (single-convolution handles)
    int res;     // conv idx whose stored output is added after the activation, or -1
    int level;   // log2 down-sampling of the OUTPUT
};

constexpr int kFd6HeadIdx = 1000;

struct Cv { int idx, cin, cout, k, s; bool bn, leaky; };

// One _conv_block (:196-215): skip source = tensor entering the second-to-last conv (:201-202).
inline int add_block(std::vector<ConvSpec>& v, const std::vector<Cv>& convs, int src, int level, bool skip) {
    int x = src, skip_src = -1;
    for (size_t c = 0; c < convs.size(); ++c) {
        if (skip && c + 2 == convs.size()) skip_src = x;
        if (convs[c].s == 2) ++level;
        const bool last = c + 1 == convs.size();
        v.push_back({convs[c].idx, convs[c].cin, convs[c].cout, convs[c].k, convs[c].s, convs[c].bn, convs[c].leaky, x,
                     (skip && last) ? skip_src : -1, level});
        x = convs[c].idx;
    }
    return x;
}

inline std::vector<ConvSpec> yolo3_table(int nb_class) {
    const int C = 3 * (5 + nb_class);
    const bool T = true, F = false;
    std::vector<ConvSpec> v;
    int x = add_block(v, {{0, 3, 32, 3, 1, T, T}, {1, 32, 64, 3, 2, T, T}, {2, 64, 32, 1, 1, T, T}, {3, 32, 64, 3, 1, T, T}}, -1, 0, true);
    x = add_block(v, {{5, 64, 128, 3, 2, T, T}, {6, 128, 64, 1, 1, T, T}, {7, 64, 128, 3, 1, T, T}}, x, 1, true);
    x = add_block(v, {{9, 128, 64, 1, 1, T, T}, {10, 64, 128, 3, 1, T, T}}, x, 2, true);
    x = add_block(v, {{12, 128, 256, 3, 2, T, T}, {13, 256, 128, 1, 1, T, T}, {14, 128, 256, 3, 1, T, T}}, x, 2, true);
    for (int i = 0; i < 7; ++i) x = add_block(v, {{16 + 3 * i, 256, 128, 1, 1, T, T}, {17 + 3 * i, 128, 256, 3, 1, T, T}}, x, 3, true);
    x = add_block(v, {{37, 256, 512, 3, 2, T, T}, {38, 512, 256, 1, 1, T, T}, {39, 256, 512, 3, 1, T, T}}, x, 3, true);
    for (int i = 0; i < 7; ++i) x = add_block(v, {{41 + 3 * i, 512, 256, 1, 1, T, T}, {42 + 3 * i, 256, 512, 3, 1, T, T}}, x, 4, true);
    x = add_block(v, {{62, 512, 1024, 3, 2, T, T}, {63, 1024, 512, 1, 1, T, T}, {64, 512, 1024, 3, 1, T, T}}, x, 4, true);
    for (int i = 0; i < 3; ++i) x = add_block(v, {{66 + 3 * i, 1024, 512, 1, 1, T, T}, {67 + 3 * i, 512, 1024, 3, 1, T, T}}, x, 5, true);
    x = add_block(v, {{75, 1024, 512, 1, 1, T, T}, {76, 512, 1024, 3, 1, T, T}, {77, 1024, 512, 1, 1, T, T},
                      {78, 512, 1024, 3, 1, T, T}, {79, 1024, 512, 1, 1, T, T}}, x, 5, false);
    add_block(v, {{80, 512, 1024, 3, 1, T, T}, {81, 1024, C, 1, 1, F, F}}, x, 5, false);
    add_block(v, {{84, 512, 256, 1, 1, T, T}}, x, 5, false);
    x = add_block(v, {{87, 768, 256, 1, 1, T, T}, {88, 256, 512, 3, 1, T, T}, {89, 512, 256, 1, 1, T, T},
                      {90, 256, 512, 3, 1, T, T}, {91, 512, 256, 1, 1, T, T}}, -2, 4, false);
    add_block(v, {{92, 256, 512, 3, 1, T, T}, {93, 512, C, 1, 1, F, F}}, x, 4, false);
    add_block(v, {{96, 256, 128, 1, 1, T, T}}, x, 4, false);
    add_block(v, {{99, 384, 128, 1, 1, T, T}, {100, 128, 256, 3, 1, T, T}, {101, 256, 128, 1, 1, T, T},
                  {102, 128, 256, 3, 1, T, T}, {103, 256, 128, 1, 1, T, T}, {104, 128, 256, 3, 1, T, T},
                  {105, 256, C, 1, 1, F, F}}, -3, 3, false);
    return v;
}

inline std::vector<ConvSpec> fd6_table(int bb_info_c_size) {
    std::vector<ConvSpec> v;
    for (const ConvSpec& c : yolo3_table(1)) if (c.idx <= 73) v.push_back(c);
    v.push_back({kFd6HeadIdx, 1024, bb_info_c_size, 3, 1, false, false, 73, -1, 5});   // face_detection.py:348-352
    return v;
}

// One convolution on its own (fvy_conv_create; stride 2: 3 x 3 reading the 4-phase form, the handle's map size is the INPUT's): no BatchNorm, no bias, no activation, input handed in by the caller
// (src = -4), dense fp32 output.
inline std::vector<ConvSpec> single_conv_table(int cin, int cout, int k, int stride) {
    return {ConvSpec{0, cin, cout, k, stride, false, false, -4, -1, stride == 2 ? 1 : 0}};
}

}  // namespace fvy
