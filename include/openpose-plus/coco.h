// COCO-18 skeleton tables of the openpose-plus public API (drop-in for the reference's
// include/openpose-plus/coco.h:5-55; values are the data contract and must match it).
#pragma once
#include <utility>
#include <vector>

constexpr int COCO_N_PARTS = 18;
constexpr int COCO_N_PAIRS = 19;

using idx_pair_t = std::pair<int, int>;
using coco_pair_list_t = std::vector<idx_pair_t>;

// limb pair_id -> (part a, part b)
const coco_pair_list_t COCOPAIRS = {{1, 2},   {1, 5},   {2, 3}, {3, 4},  {5, 6},   {6, 7},  {1, 8},
                                    {8, 9},   {9, 10},  {1, 11}, {11, 12}, {12, 13}, {1, 0},  {0, 14},
                                    {14, 16}, {0, 15},  {15, 17}, {2, 16}, {5, 17}};

// limb pair_id -> (PAF x channel, PAF y channel)
const coco_pair_list_t COCOPAIRS_NET = {{12, 13}, {20, 21}, {14, 15}, {16, 17}, {22, 23}, {24, 25}, {0, 1},
                                        {2, 3},   {4, 5},   {6, 7},   {8, 9},   {10, 11}, {28, 29}, {30, 31},
                                        {34, 35}, {32, 33}, {36, 37}, {18, 19}, {26, 27}};

// the two ear-shoulder limbs only ever join existing people
inline bool is_virtual_pair(int pair_id) { return pair_id > 16; }
