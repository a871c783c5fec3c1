// C / C++ umbrella header of openpose-plus, B200-native build (drop-in for the reference's
// include/openpose-plus.h).
#pragma once

#ifdef __cplusplus
#include <openpose-plus.hpp>
#include <openpose-plus/human.h>

extern "C" {
#endif

const int n_joins = 18 + 1;
const int n_connections = 17 + 2;

/* Declared by the reference (include/openpose-plus.h:18-22) and defined nowhere there; defined by
 * this library (csrc/opp_capi.cu): runs one frame and prints the humans found. */
extern void process_conf_paf(int height, int width, int n_joins, int n_connections,
                             const float *peaks_,  /* [n_joins, height, width] */
                             const float *pafmap_  /* [2 * n_connections, height, width] */
);

#ifdef __cplusplus
}
#endif
