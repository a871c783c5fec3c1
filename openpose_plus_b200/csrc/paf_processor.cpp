// C++ host side of the drop-in: `paf_processor` (include/openpose-plus.hpp) implemented on the C-ABI.
// Replaces the reference's paf_processor_impl class shell and factory
// (/root/reference src/paf.cpp:19-57,340-346); same constructor arguments, same call, same result type.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include <openpose-plus.hpp>

#include "../../include/opp_b200.h"

static_assert(sizeof(human_t) == sizeof(opp_human_t), "human_t must stay 292 bytes");
static_assert(sizeof(body_part_t) == sizeof(opp_body_part_t), "body_part_t must stay 16 bytes");

namespace
{
class paf_processor_b200 : public paf_batch_processor
{
  public:
    paf_processor_b200(int fh, int fw, int H, int W, int n_joins, int n_connections, int ksize, int max_batch, int device)
    {
        opp_config_default(&cfg_, fh, fw, H, W, ksize);
        cfg_.n_joins = n_joins, cfg_.n_connections = n_connections;
        cfg_.max_batch = max_batch, cfg_.device = device;
        if (opp_create(&cfg_, &h_) != OPP_OK) {
            // the reference exit(1)s on device errors (src/cudnn_traits.hpp:11-20); a constructor can throw instead
            throw std::runtime_error(std::string("create_paf_processor: ") + opp_last_error(nullptr));
        }
        size_buffers();
    }

    ~paf_processor_b200() override { opp_destroy(h_); }

    std::vector<human_t> operator()(const float *heatmap, const float *paf, bool /*use_gpu*/) override
    {
        return std::move(process_batch(heatmap, paf, 1, false)[0]);
    }

    std::vector<std::vector<human_t>> process_batch(const float *confs, const float *pafs, int n, bool device_memory) override
    {
        std::vector<std::vector<human_t>> out;
        const size_t fc = (size_t)OPP_N_HEAT * cfg_.feat_h * cfg_.feat_w, fp = (size_t)OPP_N_PAF * cfg_.feat_h * cfg_.feat_w;
        for (int done = 0; done < n;) {
            const int m = n - done < cfg_.max_batch ? n - done : cfg_.max_batch;
            opp_batch_t b;
            std::memset(&b, 0, sizeof b);
            b.conf = confs + done * fc, b.paf = pafs + done * fp, b.n_frames = m;
            b.in_mem = device_memory ? OPP_MEM_DEVICE : OPP_MEM_HOST, b.in_layout = OPP_LAYOUT_CHW, b.out_mem = OPP_MEM_HOST;
            b.humans = humans_.data(), b.n_humans = counts_.data(), b.frame_flags = flags_.data();
            if (opp_process(h_, &b) != OPP_OK) throw std::runtime_error(std::string("paf_processor: ") + opp_last_error(h_));
            // The reference's vectors grow without bound (src/post-process.h:190-198, src/paf.cpp:117-131,237-247); this
            // build works inside fixed capacities and flags a frame that exceeds one.  A caller of the reference's
            // interface must never see a truncated result: grow the capacities and run the batch again.
            int over = 0;
            for (int f = 0; f < m; ++f) over |= flags_[f] & (OPP_FLAG_PEAK_OVERFLOW | OPP_FLAG_CAND_OVERFLOW | OPP_FLAG_HUMAN_OVERFLOW);
            if (over) {
                grow(over);
                continue;
            }
            for (int f = 0; f < m; ++f) {
                std::vector<human_t> hs(counts_[f]);
                if (counts_[f]) std::memcpy(static_cast<void *>(hs.data()), &humans_[(size_t)f * cfg_.max_humans], counts_[f] * sizeof(human_t));
                out.push_back(std::move(hs));
            }
            done += m;
        }
        return out;
    }

  private:
    void size_buffers()
    {
        humans_.resize((size_t)cfg_.max_batch * cfg_.max_humans);
        counts_.resize(cfg_.max_batch), flags_.resize(cfg_.max_batch);
    }

    // doubles the capacities named by the overflow bits and re-creates the handle; throws when the device cannot hold them
    void grow(int over)
    {
        opp_config_t c = cfg_;
        if (over & OPP_FLAG_PEAK_OVERFLOW) c.max_peaks_per_part *= 2, c.max_cands_per_limb *= 4;
        if (over & OPP_FLAG_CAND_OVERFLOW) c.max_cands_per_limb *= 2;
        if (over & OPP_FLAG_HUMAN_OVERFLOW) c.max_humans *= 2;
        opp_handle_t nh = nullptr;
        if (opp_create(&c, &nh) != OPP_OK)
            throw std::runtime_error(std::string("paf_processor: frame exceeds the capacities (peaks/part ") + std::to_string(cfg_.max_peaks_per_part) +
                                     ", candidates/limb " + std::to_string(cfg_.max_cands_per_limb) + ", humans " + std::to_string(cfg_.max_humans) +
                                     ") and they cannot grow: " + opp_last_error(nullptr));
        opp_destroy(h_);
        h_ = nh, cfg_ = c;
        size_buffers();
    }

    opp_config_t cfg_;
    opp_handle_t h_ = nullptr;
    std::vector<opp_human_t> humans_;
    std::vector<int> counts_, flags_;
};
}  // namespace

paf_processor *create_paf_processor(int input_height, int input_width, int height, int width, int n_joins,
                                    int n_connections, int gauss_kernel_size)
{
    return new paf_processor_b200(input_height, input_width, height, width, n_joins, n_connections, gauss_kernel_size, 1, -1);
}

paf_batch_processor *create_paf_batch_processor(int input_height, int input_width, int height, int width,
                                                int gauss_kernel_size, int max_batch, int device)
{
    return new paf_processor_b200(input_height, input_width, height, width, 19, 19, gauss_kernel_size, max_batch, device);
}
