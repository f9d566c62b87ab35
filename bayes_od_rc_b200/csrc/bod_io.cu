// bod_io.cu — batched result writers (host code only).
//
// Reference lines replaced: the per-image tail of run_inference.py's loop,
//   run_inference.py:241-244   np.save(mean_file_name, output_boxes_vuhw)        [D,4]
//                              np.save(covar_file_name, output_covs)            [D,4,4]
//                              np.save(cat_param_file_name, output_classes)     [D,K]
//                              np.save(cat_count_file_name, output_counts)      [D,K]
//   run_inference.py:153-161   images without detections save the empty means array, shape (0,4,1), four times
// for a whole batch of padded result blocks at once, on a few host threads.  The
// files are byte-identical to what numpy.save writes (format 1.0, little-endian
// float32, C order, header padded to a multiple of 64 bytes), so the offline
// AP / MUE / PDQ scripts load them unchanged.
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bayesod.h"

namespace {

// numpy.lib.format.write_array_header_1_0 for '<f4', C order
std::string npy_header(const int* shape, int ndim) {
    std::string d = "{'descr': '<f4', 'fortran_order': False, 'shape': (";
    for (int i = 0; i < ndim; ++i) {
        d += std::to_string(shape[i]);
        if (ndim == 1) d += ",";
        else if (i + 1 < ndim) d += ", ";
    }
    d += "), }";
    // magic (6) + version (2) + header length (2) + header, '\n'-terminated, padded with spaces to a multiple of 64
    const size_t unpadded = 10 + d.size() + 1;
    const size_t pad = (64 - unpadded % 64) % 64;
    d.append(pad, ' ');
    d += '\n';
    std::string out("\x93NUMPY\x01\x00", 8);
    const uint16_t hl = (uint16_t)d.size();
    out += (char)(hl & 0xff);
    out += (char)(hl >> 8);
    out += d;
    return out;
}

bool write_npy(const std::string& path, const float* data, const int* shape, int ndim) {
    size_t n = 1;
    for (int i = 0; i < ndim; ++i) n *= (size_t)shape[i];
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const std::string h = npy_header(shape, ndim);
    bool ok = fwrite(h.data(), 1, h.size(), f) == h.size();
    if (ok && n) ok = fwrite(data, sizeof(float), n, f) == n;
    return (fclose(f) == 0) && ok;
}

}  // namespace

extern "C" int bod_write_npy(const char* path, const float* data, const int32_t* shape, int32_t ndim) {
    if (!path || !shape || ndim < 1 || ndim > 8) return BOD_ERR_INVALID;
    int sh[8];
    for (int i = 0; i < ndim; ++i) { if (shape[i] < 0) return BOD_ERR_INVALID; sh[i] = shape[i]; }
    return write_npy(path, data, sh, ndim) ? BOD_OK : BOD_ERR_STATE;
}

extern "C" int bod_write_results_npy(const bod_host_results* res, int32_t B, int32_t Dmax, int32_t K,
                                     const char* mean_dir, const char* cov_dir, const char* cat_param_dir,
                                     const char* cat_count_dir, const char* const* sample_ids, int32_t nthreads) {
    if (!res || !res->num_dets || !res->means || !res->covs || !res->cat_param || !res->cat_count || !sample_ids ||
        !mean_dir || !cov_dir || !cat_param_dir || !cat_count_dir || B < 1 || Dmax < 1 || K < 1)
        return BOD_ERR_INVALID;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > B) nthreads = B;
    std::vector<int> failed((size_t)nthreads, 0);
    auto work = [&](int t) {
        for (int b = t; b < B; b += nthreads) {
            const int D = res->num_dets[b];
            const std::string id = std::string(sample_ids[b]) + ".npy";
            bool ok = true;
            if (D <= 0) {                                   // run_inference.py:153-161
                const int sh[3] = {0, 4, 1};
                for (const char* dir : {mean_dir, cov_dir, cat_param_dir, cat_count_dir})
                    ok = write_npy(std::string(dir) + "/" + id, nullptr, sh, 3) && ok;
            } else {
                const int s_mean[2] = {D, 4}, s_cov[3] = {D, 4, 4}, s_cat[2] = {D, K};
                ok = write_npy(std::string(mean_dir) + "/" + id, res->means + (size_t)b * Dmax * 4, s_mean, 2) && ok;
                ok = write_npy(std::string(cov_dir) + "/" + id, res->covs + (size_t)b * Dmax * 16, s_cov, 3) && ok;
                ok = write_npy(std::string(cat_param_dir) + "/" + id, res->cat_param + (size_t)b * Dmax * K, s_cat, 2) && ok;
                ok = write_npy(std::string(cat_count_dir) + "/" + id, res->cat_count + (size_t)b * Dmax * K, s_cat, 2) && ok;
            }
            if (!ok) failed[(size_t)t] = 1;
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    for (int f : failed) if (f) return BOD_ERR_STATE;
    return BOD_OK;
}
