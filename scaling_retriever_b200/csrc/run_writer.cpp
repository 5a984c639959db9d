// Result materialisation on the host: run.json straight from the [Q, k] result arrays.
//
// Replaces the per-pair work of the reference after the search — `res[str(qid)][str(doc_ids[id_])] = float(sc)`
// (scaling_retriever/indexer.py:429-430, eval_dense.py:229-241: Q*k str()/dict operations under the GIL) and
// `json.dump(res, handler)` (indexer.py:537-538) — which takes ~10 s for 6,980 x 1000 pairs around a 0.11 s search.
// The output is BYTE-IDENTICAL to `json.dumps(res)` of the dict the reference builds (CPython json with default
// arguments: ", " / ": " separators, ensure_ascii=True, float.__repr__ for the scores), provided that the query ids are
// distinct and the external ids of one row are distinct (the Python caller checks both and otherwise takes the dict path).
//
// Host code only (no device work): queries are formatted in parallel (OpenMP) into per-chunk buffers and written in order.
#include <omp.h>

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200ret.h"

namespace b200ret {
void set_err(const char* fmt, ...);
}

namespace {

// json.encoder.py_encode_basestring_ascii: '"' + escaped + '"'; input is UTF-8.
void append_json_string(std::string& out, const char* s, size_t n) {
    static const char HEX[] = "0123456789abcdef";
    out.push_back('"');
    const unsigned char* p = reinterpret_cast<const unsigned char*>(s);
    const unsigned char* end = p + n;
    auto put_u = [&](unsigned cp) {
        out += "\\u";
        out.push_back(HEX[(cp >> 12) & 15]);
        out.push_back(HEX[(cp >> 8) & 15]);
        out.push_back(HEX[(cp >> 4) & 15]);
        out.push_back(HEX[cp & 15]);
    };
    while (p < end) {
        unsigned c = *p;
        if (c >= 0x20 && c <= 0x7e && c != '"' && c != '\\') {
            out.push_back(static_cast<char>(c));
            ++p;
            continue;
        }
        if (c < 0x80) {
            switch (c) {
                case '"': out += "\\\""; break;
                case '\\': out += "\\\\"; break;
                case '\n': out += "\\n"; break;
                case '\r': out += "\\r"; break;
                case '\t': out += "\\t"; break;
                case '\b': out += "\\b"; break;
                case '\f': out += "\\f"; break;
                default: put_u(c);
            }
            ++p;
            continue;
        }
        // decode one UTF-8 sequence (the caller encodes Python str with 'utf-8', so the input is well formed)
        unsigned cp = 0;
        int extra = 0;
        if ((c & 0xE0) == 0xC0) { cp = c & 0x1F; extra = 1; }
        else if ((c & 0xF0) == 0xE0) { cp = c & 0x0F; extra = 2; }
        else if ((c & 0xF8) == 0xF0) { cp = c & 0x07; extra = 3; }
        else { cp = 0xFFFD; extra = 0; }
        ++p;
        for (int i = 0; i < extra && p < end; ++i, ++p) cp = (cp << 6) | (*p & 0x3F);
        if (cp >= 0x10000) {
            cp -= 0x10000;
            put_u(0xD800 | (cp >> 10));
            put_u(0xDC00 | (cp & 0x3FF));
        } else {
            put_u(cp);
        }
    }
    out.push_back('"');
}

// float.__repr__ of the double value of an fp32 score (what json.dumps writes for float(np.float32)):
// shortest round-trip digits; exponent form iff decimal exponent < -4 or >= 16; ".0" appended to integers.
void append_py_float(std::string& out, float f) {
    const double v = static_cast<double>(f);
    if (std::isnan(v)) { out += "NaN"; return; }
    if (std::isinf(v)) { out += (v < 0 ? "-Infinity" : "Infinity"); return; }
    if (v == 0.0) { out += (std::signbit(v) ? "-0.0" : "0.0"); return; }
    char buf[40];
    auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);   // shortest: "[-]d[.ddd]e[+-]XX"
    const char* p = buf;
    const char* end = r.ptr;
    if (*p == '-') { out.push_back('-'); ++p; }
    char digits[24];
    int nd = 0;
    while (p < end && *p != 'e') {
        if (*p != '.') digits[nd++] = *p;
        ++p;
    }
    ++p;   // 'e'
    int esign = 1;
    if (*p == '-') { esign = -1; ++p; } else if (*p == '+') { ++p; }
    int e10 = 0;
    while (p < end) e10 = e10 * 10 + (*p++ - '0');
    e10 *= esign;
    const int decpt = e10 + 1;   // value = 0.d1d2... * 10^decpt
    if (decpt <= -4 || decpt > 16) {
        out.push_back(digits[0]);
        if (nd > 1) {
            out.push_back('.');
            out.append(digits + 1, nd - 1);
        }
        out.push_back('e');
        out.push_back(e10 < 0 ? '-' : '+');
        const int a = e10 < 0 ? -e10 : e10;
        if (a < 10) out.push_back('0');
        char eb[8];
        auto er = std::to_chars(eb, eb + sizeof(eb), a);
        out.append(eb, er.ptr - eb);
    } else if (decpt <= 0) {
        out += "0.";
        out.append(static_cast<size_t>(-decpt), '0');
        out.append(digits, nd);
    } else if (decpt >= nd) {
        out.append(digits, nd);
        out.append(static_cast<size_t>(decpt - nd), '0');
        out += ".0";
    } else {
        out.append(digits, decpt);
        out.push_back('.');
        out.append(digits + decpt, nd - decpt);
    }
}

void append_int(std::string& out, int64_t v) {
    char buf[24];
    auto r = std::to_chars(buf, buf + sizeof(buf), v);
    out.append(buf, r.ptr - buf);
}

}  // namespace

extern "C" int b200ret_write_run_json(const char* path_host, const int64_t* ids_host, const float* scores_host,
                                      const int32_t* counts_host, int32_t n_queries, int32_t row_stride,
                                      const char* qid_blob_host, const int64_t* qid_offsets_host,
                                      const char* docid_blob_host, const int64_t* docid_offsets_host,
                                      const int64_t* docid_ints_host, int64_t n_doc_ids, int64_t* bytes_written_host) {
    using b200ret::set_err;
    if (!path_host || n_queries < 0 || row_stride < 0 || (n_queries > 0 && (!ids_host || !scores_host || !qid_blob_host || !qid_offsets_host))) {
        set_err("write_run_json: null pointer / negative size");
        return B200RET_EINVAL;
    }
    if ((docid_blob_host != nullptr) != (docid_offsets_host != nullptr) || (docid_blob_host && docid_ints_host)) {
        set_err("write_run_json: pass either the string table (blob + offsets) or the integer table, not both");
        return B200RET_EINVAL;
    }
    // validate once, serially: row labels inside the table, counts inside the row
    for (int32_t q = 0; q < n_queries; ++q) {
        const int32_t c = counts_host ? counts_host[q] : row_stride;
        if (c < 0 || c > row_stride) {
            set_err("write_run_json: counts[%d]=%d outside [0, %d]", q, c, row_stride);
            return B200RET_EINVAL;
        }
    }
    const int n_chunks = n_queries > 0 ? std::min(n_queries, 256) : 0;
    std::vector<std::string> chunks(static_cast<size_t>(n_chunks));
    std::vector<int> chunk_live(static_cast<size_t>(n_chunks), 0);
    int bad = 0;
    // explicit team size: torch.distributed.run exports OMP_NUM_THREADS=1 to every rank, and only the first worker formats a run
    const int n_threads = std::max(1, std::min(omp_get_num_procs(), 32));
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
    for (int ch = 0; ch < n_chunks; ++ch) {
        const int32_t q0 = static_cast<int32_t>(static_cast<int64_t>(n_queries) * ch / n_chunks);
        const int32_t q1 = static_cast<int32_t>(static_cast<int64_t>(n_queries) * (ch + 1) / n_chunks);
        std::string& out = chunks[ch];
        out.reserve(static_cast<size_t>(q1 - q0) * (static_cast<size_t>(row_stride) * 32 + 32));
        for (int32_t q = q0; q < q1; ++q) {
            const int32_t c = counts_host ? counts_host[q] : row_stride;
            if (c == 0) continue;   // a query without an eligible doc has NO key (defaultdict semantics, indexer.py:415,429)
            if (chunk_live[ch]++) out += ", ";
            append_json_string(out, qid_blob_host + qid_offsets_host[q], static_cast<size_t>(qid_offsets_host[q + 1] - qid_offsets_host[q] - 1));
            out += ": {";
            const int64_t* row = ids_host + static_cast<size_t>(q) * row_stride;
            const float* sc = scores_host + static_cast<size_t>(q) * row_stride;
            for (int32_t j = 0; j < c; ++j) {
                int64_t id = row[j];
                if (n_doc_ids > 0 && id < 0) id += n_doc_ids;   // Python negative indexing (faiss -1 label, indexer.py:212)
                if (j) out += ", ";
                if (docid_blob_host) {
                    if (id < 0 || id >= n_doc_ids) {
#pragma omp atomic write
                        bad = 1;
                        id = 0;
                    }
                    append_json_string(out, docid_blob_host + docid_offsets_host[id],
                                       static_cast<size_t>(docid_offsets_host[id + 1] - docid_offsets_host[id] - 1));
                } else {
                    if (docid_ints_host) {
                        if (id < 0 || id >= n_doc_ids) {
#pragma omp atomic write
                            bad = 1;
                            id = 0;
                        }
                        id = docid_ints_host[id];
                    }
                    out.push_back('"');       // str(int): digits only, nothing to escape
                    append_int(out, id);
                    out.push_back('"');
                }
                out += ": ";
                append_py_float(out, sc[j]);
            }
            out.push_back('}');
        }
    }
    if (bad) {
        set_err("write_run_json: a row label is outside the external-id table (%lld entries)", (long long)n_doc_ids);
        return B200RET_EINVAL;
    }
    FILE* f = fopen(path_host, "wb");
    if (!f) {
        set_err("write_run_json: cannot open %s", path_host);
        return B200RET_EINVAL;
    }
    int64_t total = 0;
    bool ok = fputc('{', f) != EOF;
    total += 1;
    bool first = true;
    for (int ch = 0; ch < n_chunks && ok; ++ch) {
        if (!chunk_live[ch]) continue;
        if (!first) { ok = fwrite(", ", 1, 2, f) == 2; total += 2; }
        first = false;
        ok = ok && fwrite(chunks[ch].data(), 1, chunks[ch].size(), f) == chunks[ch].size();
        total += static_cast<int64_t>(chunks[ch].size());
    }
    ok = ok && fputc('}', f) != EOF;
    total += 1;
    ok = (fclose(f) == 0) && ok;
    if (!ok) {
        set_err("write_run_json: write to %s failed", path_host);
        return B200RET_EINVAL;
    }
    if (bytes_written_host) *bytes_written_host = total;
    return B200RET_OK;
}

// Parallel host memcpy (result rows out of the reusable pinned staging buffers into arrays the caller owns): a fresh 84 MB
// destination is page-faulted in by several threads instead of one.
extern "C" int b200ret_host_copy(void* dst_host, const void* src_host, size_t bytes, int n_threads) {
    if (bytes == 0) return B200RET_OK;
    if (!dst_host || !src_host) {
        b200ret::set_err("host_copy: null pointer");
        return B200RET_EINVAL;
    }
    const int n = std::max(1, std::min({n_threads > 0 ? n_threads : 8, omp_get_num_procs(), static_cast<int>(bytes / (1 << 20)) + 1}));
    const size_t part = (bytes / n + 4095) & ~static_cast<size_t>(4095);
#pragma omp parallel for schedule(static, 1) num_threads(n)
    for (int t = 0; t < n; ++t) {
        const size_t a = std::min(bytes, part * t), b = std::min(bytes, part * (t + 1));
        if (b > a) memcpy(static_cast<char*>(dst_host) + a, static_cast<const char*>(src_host) + a, b - a);
    }
    return B200RET_OK;
}
