// Host-side batch formatter and subset aggregator for the result files either side of the hot
// path.  Pure C++ (no CUDA): once the distances come off the GPU at
// millions of pairs per second, the reference's per-value `str.format` + generator `.send`
// writers (distances.py:95-186, versus_all.py:278-350) and its per-pair dict aggregation
// (versus_all.py:57-95, 623-645) would dominate the task by orders of magnitude.
//
// Byte-for-byte contract: rows are "\t".join(fields) + "\n" (handlers.py:219-227); floats are
// printf-formatted with the C equivalent of the Python format spec (both are correctly rounded, so
// "{:.4f}" and "%.4f" agree); undefined values print the `missing` marker.  The aggregation adds
// the values in the reference's row-major order, so the fp64 sums are bit-identical.
#include "../../include/taxi2_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Table {
    const char* bytes;
    const int64_t* off;
    void append(std::string& out, int64_t k) const { out.append(bytes + off[k], (size_t)(off[k + 1] - off[k])); }
};

inline void append_value(std::string& out, double v, bool undefined, double scale, const char* fmt, const char* missing)
{
    if (undefined || std::isnan(v) || std::isinf(v)) { out += missing; return; }
    char buf[64];
    const int n = std::snprintf(buf, sizeof buf, fmt, v * scale);
    out.append(buf, (size_t)std::max(0, std::min<int>(n, (int)sizeof buf - 1)));
}

// run fn(row_begin, row_end, out) over row chunks on `threads` threads, then write the chunks in order
template <class Fn> int format_rows(const char* path, int32_t nx, int32_t threads, Fn fn)
{
    if (threads <= 0) threads = (int32_t)std::max(1u, std::thread::hardware_concurrency());
    threads = std::max(1, std::min<int32_t>(threads, std::max(1, nx)));
    std::vector<std::string> chunks((size_t)threads);
    std::vector<std::thread> pool;
    const int32_t per = (nx + threads - 1) / threads;
    for (int32_t t = 0; t < threads; ++t) {
        const int32_t b = t * per, e = std::min(nx, b + per);
        if (b >= e) break;
        if (threads == 1) fn(b, e, chunks[(size_t)t]);
        else pool.emplace_back([&, b, e, t] { fn(b, e, chunks[(size_t)t]); });
    }
    for (auto& th : pool) th.join();
    FILE* f = std::fopen(path, "ab");
    if (!f) return TAXI_E_ARG;
    for (const auto& c : chunks)
        if (!c.empty() && std::fwrite(c.data(), 1, c.size(), f) != c.size()) { std::fclose(f); return TAXI_E_ARG; }
    return std::fclose(f) == 0 ? TAXI_OK : TAXI_E_ARG;
}

}  // namespace

extern "C" {

// Row layout segments of taxi_format_pairs
enum { SEG_X0 = 0, SEG_X1, SEG_X2, SEG_X3, SEG_Y0, SEG_Y1, SEG_Y2, SEG_Y3, SEG_SCORES, SEG_COMPARISON };

int taxi_format_pairs(const char* path, const int32_t* segments, int32_t nsegments,
                      const char* const* xbytes, const int64_t* const* xoff,
                      const char* const* ybytes, const int64_t* const* yoff,
                      int32_t x0, int32_t nx, int32_t ny,
                      const double* metrics, const uint8_t* undefined,
                      const int32_t* columns, int32_t ncolumns, double scale,
                      const char* float_format, const char* missing,
                      const int32_t* xgenus, const int32_t* xspecies, const int32_t* ygenus, const int32_t* yspecies,
                      const char* const* type_labels, int32_t threads)
{
    if (!path || !segments || nsegments <= 0 || nx < 0 || ny < 0 || !metrics || !float_format || !missing) return TAXI_E_ARG;
    auto fn = [&](int32_t b, int32_t e, std::string& out) {
        out.reserve((size_t)(e - b) * (size_t)ny * 48);
        for (int32_t i = b; i < e; ++i) {
            for (int32_t j = 0; j < ny; ++j) {
                const size_t p = (size_t)i * ny + j;
                const bool undef = undefined && undefined[p];
                for (int32_t s = 0; s < nsegments; ++s) {
                    if (s) out += '\t';
                    const int seg = segments[s];
                    if (seg >= SEG_X0 && seg <= SEG_X3) Table{xbytes[seg], xoff[seg]}.append(out, x0 + i);
                    else if (seg >= SEG_Y0 && seg <= SEG_Y3) Table{ybytes[seg - SEG_Y0], yoff[seg - SEG_Y0]}.append(out, j);
                    else if (seg == SEG_SCORES) {
                        for (int32_t c = 0; c < ncolumns; ++c) {
                            if (c) out += '\t';
                            append_value(out, metrics[p * 4 + columns[c]], undef, scale, float_format, missing);
                        }
                    } else if (seg == SEG_COMPARISON) {
                        // versus_all.py:262-275: None when a partition is absent, else "same subset?"
                        const int sg = xgenus ? (xgenus[x0 + i] == ygenus[j] ? 1 : 0) : -1;
                        const int ss = xspecies ? (xspecies[x0 + i] == yspecies[j] ? 1 : 0) : -1;
                        int label;   // 0 Unknown, 1 IntraSpecies, 2 InterSpecies, 3 IntraGenus, 4 InterGenus
                        if (sg < 0) label = ss < 0 ? 0 : (ss ? 1 : 2);
                        else if (sg == 0) label = 4;
                        else label = ss < 0 ? 3 : (ss ? 1 : 2);
                        out += type_labels[label];
                    }
                }
                out += '\n';
            }
        }
    };
    return format_rows(path, nx, threads, fn);
}

// One row per x: id, then one value per y (DistanceHandler.Matrix rows, distances.py:183-186)
int taxi_format_matrix(const char* path, const char* xid_bytes, const int64_t* xid_off, int32_t x0, int32_t nx, int32_t ny,
                       const double* metrics, const uint8_t* undefined, int32_t column, double scale,
                       const char* float_format, const char* missing, int32_t threads)
{
    if (!path || !xid_bytes || !xid_off || nx < 0 || ny < 0 || !metrics || column < 0 || column > 3) return TAXI_E_ARG;
    const Table ids{xid_bytes, xid_off};
    auto fn = [&](int32_t b, int32_t e, std::string& out) {
        out.reserve((size_t)(e - b) * (size_t)ny * 8);
        for (int32_t i = b; i < e; ++i) {
            ids.append(out, x0 + i);
            for (int32_t j = 0; j < ny; ++j) {
                const size_t p = (size_t)i * ny + j;
                out += '\t';
                append_value(out, metrics[p * 4 + column], undefined && undefined[p], scale, float_format, missing);
            }
            out += '\n';
        }
    };
    return format_rows(path, nx, threads, fn);
}

// SequencePairHandler.Formatted records (pairs.py:51-97) for the nx*ny pairs of a row block, pair
// p = (x0 + p / ny, p % ny): "idx / idy", aligned x, the match pattern ('|' equal and not a gap, '-'
// either is a gap, '.' otherwise), aligned y; records are separated by one empty line, so every
// record but the file's first is preceded by "\n".  The gapped strings are the right-aligned
// slots taxi_align_strings fills: pair p occupies [aln_start[p], aln_off[p + 1]).
int taxi_format_aligned_pairs(const char* path, int32_t first_record,
                              const char* xid_bytes, const int64_t* xid_off, const char* yid_bytes, const int64_t* yid_off,
                              int32_t x0, int32_t nx, int32_t ny,
                              const uint8_t* aln_x, const uint8_t* aln_y, const int64_t* aln_start, const int64_t* aln_off,
                              int32_t threads)
{
    if (!path || !xid_bytes || !xid_off || !yid_bytes || !yid_off || nx < 0 || ny < 0 || !aln_x || !aln_y || !aln_start || !aln_off)
        return TAXI_E_ARG;
    const Table xids{xid_bytes, xid_off}, yids{yid_bytes, yid_off};
    auto fn = [&](int32_t b, int32_t e, std::string& out) {
        size_t need = 0;
        for (int64_t p = (int64_t)b * ny; p < (int64_t)e * ny; ++p) need += 3 * (size_t)(aln_off[p + 1] - aln_start[p]) + 48;
        out.reserve(need);
        for (int32_t i = b; i < e; ++i) {
            for (int32_t j = 0; j < ny; ++j) {
                const int64_t p = (int64_t)i * ny + j;
                if (p > 0 || !first_record) out += '\n';
                xids.append(out, x0 + i);
                out += " / ";
                yids.append(out, j);
                out += '\n';
                const uint8_t* ax = aln_x + aln_start[p];
                const uint8_t* ay = aln_y + aln_start[p];
                const size_t len = (size_t)(aln_off[p + 1] - aln_start[p]);
                out.append(reinterpret_cast<const char*>(ax), len);
                out += '\n';
                const size_t at = out.size();
                out.resize(at + len);
                for (size_t k = 0; k < len; ++k) {
                    const uint8_t a = ax[k], c = ay[k];
                    out[at + k] = (a == '-' || c == '-') ? '-' : (a == c ? '|' : '.');
                }
                out += '\n';
                out.append(reinterpret_cast<const char*>(ay), len);
                out += '\n';
            }
        }
    };
    return format_rows(path, nx, threads, fn);
}

// SimpleAggregator state per (subset_x, subset_y) for one metric column (versus_all.py:57-95):
// sum / min / max / n over the defined values in row-major order, plus the order in which the keys
// first appeared (the reference's dicts iterate in insertion order).  Subset ids are 0..nsub-1, with
// one id reserved by the caller for "not in the partition".
int taxi_aggregate_subsets(const double* metrics, const uint8_t* undefined, int32_t x0, int32_t nx, int32_t ny,
                           int32_t column, double scale, const int32_t* xsubset, const int32_t* ysubset, int32_t nsub,
                           double* sum, double* vmin, double* vmax, int64_t* count, int64_t* first_seen, int64_t* next_order)
{
    if (!metrics || !xsubset || !ysubset || !sum || !vmin || !vmax || !count || !first_seen || !next_order) return TAXI_E_ARG;
    for (int32_t i = 0; i < nx; ++i) {
        const int64_t sx = xsubset[x0 + i];
        for (int32_t j = 0; j < ny; ++j) {
            const size_t p = (size_t)i * ny + j;
            const int64_t k = sx * nsub + ysubset[j];
            if (first_seen[k] < 0) first_seen[k] = (*next_order)++;
            double v = metrics[p * 4 + column];
            if ((undefined && undefined[p]) || std::isnan(v) || std::isinf(v)) continue;
            v *= scale;
            sum[k] += v;
            if (v < vmin[k]) vmin[k] = v;
            if (v > vmax[k]) vmax[k] = v;
            count[k] += 1;
        }
    }
    return TAXI_OK;
}

}  // extern "C"
