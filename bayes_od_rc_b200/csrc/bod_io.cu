// bod_io.cu — batched result writers (host code only):
// the four .npy files per image (below), the BDD predictions.json and the KITTI
// text files (second half of this file).
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
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
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

// ---------------------------------------------------------------------------------------
// Text writers.  Reference lines replaced:
//   validation_utils.py:183-213  predictions_to_bdd_format   (one dict per detection)
//   run_inference.py:206-212     final_results_list.extend(...) per image
//   run_inference.py:258-260     json.dump(final_results_list, fp, indent=4, separators=(',', ': '))
//   validation_utils.py:216-272  predictions_to_kitti_format (rows of strings)
//   run_inference.py:176-201     np.savetxt(<id>.txt, rows, newline='\r\n', fmt='%s') / np.savetxt(<id>.txt, [])
//   box_utils.py:70-88           vuhw_to_vuvu_np (float32: v -+ h / 2.0)
// Byte-identical output needs the two number formats involved: Python's float.__repr__
// of the binary64 value (json) and numpy's str() of a float32 scalar (the KITTI rows
// become a '<U32' array).  Both are "shortest digits that round-trip" (std::to_chars)
// laid out by slightly different rules.
namespace {

struct Digits { bool neg; std::string d; int exp10; };       // value = 0.d[0] d[1].. x 10^(exp10+1) -> d[0].d[1..] x 10^exp10

template <class T> Digits shortest_digits(T v) {
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf - 1, v, std::chars_format::scientific);
    *r.ptr = '\0';
    Digits out{false, "", 0};
    const char* p = buf;
    if (*p == '-') { out.neg = true; ++p; }
    for (; p < r.ptr && *p != 'e'; ++p) if (*p != '.') out.d += *p;
    out.exp10 = (int)strtol(p + 1, nullptr, 10);
    return out;
}

std::string layout(const Digits& g, bool scientific) {
    std::string s = g.neg ? "-" : "";
    const int n = (int)g.d.size();
    if (scientific) {
        s += g.d[0];
        if (n > 1) { s += '.'; s.append(g.d, 1, std::string::npos); }
        const int e = g.exp10 < 0 ? -g.exp10 : g.exp10;
        s += 'e'; s += g.exp10 < 0 ? '-' : '+';
        if (e < 10) s += '0';
        s += std::to_string(e);
    } else if (g.exp10 >= 0) {
        for (int i = 0; i <= g.exp10; ++i) s += i < n ? g.d[(size_t)i] : '0';
        s += '.';
        if (n > g.exp10 + 1) s.append(g.d, (size_t)g.exp10 + 1, std::string::npos); else s += '0';
    } else {
        s += "0.";
        s.append((size_t)(-g.exp10 - 1), '0');
        s += g.d;
    }
    return s;
}

// float.__repr__ (Python >= 3.1, 'r' format): exponent form iff decpt > 16 or decpt < -3
std::string py_float_repr(double v) {
    if (std::isnan(v)) return "NaN";                          // json.dumps(allow_nan=True) spellings
    if (std::isinf(v)) return v < 0 ? "-Infinity" : "Infinity";
    const Digits g = shortest_digits(v);
    return layout(g, g.exp10 >= 16 || g.exp10 < -4);
}

// str(numpy.float32): positional iff 0 or 1e-4 <= |x| < 1e6 (on the exact value; the upper limit is 1e16 for
// float64 only), else exponent form.  np.asarray of the mixed KITTI rows converts its float32 cells the same way.
std::string np_float32_str(float v) {
    if (std::isnan(v)) return "nan";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    const double a = std::fabs((double)v);
    return layout(shortest_digits(v), !(a == 0.0 || (a < 1e6 && a >= 1e-4)));
}

// json.encoder.py_encode_basestring_ascii
std::string json_string(const char* z) {
    std::string s = "\"";
    const unsigned char* p = (const unsigned char*)z;
    auto u4 = [&](unsigned c) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); s += b; };
    while (*p) {
        unsigned c = *p++;
        if (c == '"') s += "\\\"";
        else if (c == '\\') s += "\\\\";
        else if (c == '\n') s += "\\n";
        else if (c == '\r') s += "\\r";
        else if (c == '\t') s += "\\t";
        else if (c == '\b') s += "\\b";
        else if (c == '\f') s += "\\f";
        else if (c >= 0x20 && c < 0x7f) s += (char)c;
        else if (c < 0x80) u4(c);
        else {                                               // UTF-8 -> code point -> \uXXXX (surrogate pair above the BMP)
            int extra = c >= 0xf0 ? 3 : c >= 0xe0 ? 2 : 1;
            unsigned cp = c & (0x3f >> extra);
            while (extra-- && (*p & 0xc0) == 0x80) cp = (cp << 6) | (*p++ & 0x3f);
            if (cp >= 0x10000) { cp -= 0x10000; u4(0xd800 | (cp >> 10)); u4(0xdc00 | (cp & 0x3ff)); }
            else u4(cp);
        }
    }
    return s + "\"";
}

// numpy.argmax of one class row: first maximum, a NaN wins over everything after it
int first_argmax(const float* row, int K) {
    int m = 0;
    for (int k = 1; k < K; ++k) {
        if (std::isnan(row[m])) break;
        if (row[k] > row[m] || std::isnan(row[k])) m = k;
    }
    return m;
}

// box_utils.py:70-88 on one float32 row (numpy keeps float32 when dividing by the Python scalar 2.0)
void vuhw_to_vuvu_row(const float* m, float out[4]) {
    const float hh = m[2] / 2.0f, hw = m[3] / 2.0f;
    out[0] = m[0] - hh; out[1] = m[1] - hw; out[2] = m[0] + hh; out[3] = m[1] + hw;
}

bool results_ok(const bod_host_results* res, int32_t B, int32_t Dmax, int32_t K, const char* const* ids) {
    return res && res->num_dets && res->means && res->cat_param && ids && B >= 1 && Dmax >= 1 && K >= 1;
}

}  // namespace

struct bod_json_writer {
    FILE* f = nullptr;
    std::vector<std::string> categories;
    size_t entries = 0;
    bool failed = false;
};

extern "C" int bod_bdd_json_open(bod_json_writer** out, const char* path, const char* const* categories,
                                 int32_t n_categories) {
    if (!out || !path || n_categories < 0 || (n_categories && !categories)) return BOD_ERR_INVALID;
    FILE* f = fopen(path, "wb");
    if (!f) return BOD_ERR_STATE;
    auto* w = new bod_json_writer;
    w->f = f;
    for (int i = 0; i < n_categories; ++i) w->categories.push_back(json_string(categories[i]));
    *out = w;
    return BOD_OK;
}

extern "C" int bod_bdd_json_append(bod_json_writer* w, const bod_host_results* res, int32_t B, int32_t Dmax, int32_t K,
                                   const char* const* sample_ids) {
    if (!w || !w->f || !results_ok(res, B, Dmax, K, sample_ids)) return BOD_ERR_INVALID;
    std::string s;
    for (int b = 0; b < B; ++b) {
        const int D = res->num_dets[b] < Dmax ? res->num_dets[b] : Dmax;
        const std::string name = json_string(sample_ids[b]);
        for (int d = 0; d < D; ++d) {
            const float* cls = res->cat_param + ((size_t)b * Dmax + d) * K;
            const int m = first_argmax(cls, K);
            if (m >= (int)w->categories.size()) continue;                    // validation_utils.py:203
            float c[4];
            vuhw_to_vuvu_row(res->means + ((size_t)b * Dmax + d) * 4, c);
            s += w->entries++ ? ",\n    {\n" : "[\n    {\n";
            s += "        \"name\": " + name + ",\n        \"timestep\": 1000,\n        \"category\": " + w->categories[(size_t)m] +
                 ",\n        \"bbox\": [\n";
            const float bbox[4] = {c[1], c[0], c[3], c[2]};                  // [u_min, v_min, u_max, v_max], :206-209
            for (int i = 0; i < 4; ++i) s += "            " + py_float_repr((double)bbox[i]) + (i < 3 ? ",\n" : "\n");
            s += "        ],\n        \"score\": " + py_float_repr((double)cls[m]) + "\n    }";
        }
    }
    if (!s.empty() && fwrite(s.data(), 1, s.size(), w->f) != s.size()) { w->failed = true; return BOD_ERR_STATE; }
    return BOD_OK;
}

extern "C" int bod_bdd_json_close(bod_json_writer* w) {
    if (!w) return BOD_ERR_INVALID;
    bool ok = !w->failed;
    if (w->f) {
        const char* tail = w->entries ? "\n]" : "[]";
        ok = fwrite(tail, 1, strlen(tail), w->f) == strlen(tail) && ok;
        ok = fclose(w->f) == 0 && ok;
    }
    delete w;
    return ok ? BOD_OK : BOD_ERR_STATE;
}

extern "C" int bod_write_results_kitti_txt(const bod_host_results* res, int32_t B, int32_t Dmax, int32_t K,
                                           const char* dir, const char* const* sample_ids, int32_t nthreads) {
    if (!results_ok(res, B, Dmax, K, sample_ids) || !dir) return BOD_ERR_INVALID;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > B) nthreads = B;
    std::vector<int> failed((size_t)nthreads, 0);
    static const char* const kNames[2] = {"Car", "Pedestrian"};              // validation_utils.py:235-252
    auto work = [&](int t) {
        std::string s;
        for (int b = t; b < B; b += nthreads) {
            s.clear();
            const int D = res->num_dets[b] < Dmax ? res->num_dets[b] : Dmax;
            for (int d = 0; d < D; ++d) {
                const float* cls = res->cat_param + ((size_t)b * Dmax + d) * K;
                const int m = first_argmax(cls, K);
                if (m > 1) continue;
                float c[4];
                vuhw_to_vuvu_row(res->means + ((size_t)b * Dmax + d) * 4, c);
                s += kNames[m];
                s += " -1 -1 -10 " + np_float32_str(c[1]) + " " + np_float32_str(c[0]) + " " + np_float32_str(c[3]) + " " +
                     np_float32_str(c[2]) + " -10 -10 -10 -10 -10 -10 -10 " + np_float32_str(cls[m]) + "\r\n";
            }
            FILE* f = fopen((std::string(dir) + "/" + sample_ids[b] + ".txt").c_str(), "wb");
            bool ok = f != nullptr;
            if (ok && !s.empty()) ok = fwrite(s.data(), 1, s.size(), f) == s.size();
            if (f) ok = fclose(f) == 0 && ok;
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

extern "C" int bod_format_float(double value, int32_t style, char* out, int32_t cap) {
    if (!out || cap < 1) return BOD_ERR_INVALID;
    const std::string s = style == 0 ? py_float_repr(value) : np_float32_str((float)value);
    if ((int)s.size() + 1 > cap) return BOD_ERR_INVALID;
    memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}
