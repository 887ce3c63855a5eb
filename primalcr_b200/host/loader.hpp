// loader.hpp -- fast host loader for the reference's data directory format (replaces the serial fgets+sscanf+std::sort
// of load() util.cpp:6-25 / smat_t::load_from_iterator util.h:201-271 / testset_t::load util.h:360-371).
//
//   data_dir/meta : "m n" / "nnz_train train_file" / optional "nnz_test test_file"        (util.cpp:9-21)
//   ratings files : "user item rating" per line, 1-based ids                               (util.h:126, 367)
//
// The file is mmap'ed, split at line boundaries into one chunk per thread and parsed with hand-rolled integer / strtod
// scanners; training entries are bucketed by user (counting sort) and ordered by item inside a user, which is the order
// smat_t's `sort(perm, SparseComp)` (util.h:240) produces; the test set keeps file order inside a user and must be
// grouped by user like convert(testset_t&) requires (util.cpp:257-266).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <stdexcept>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace pcrhost {

struct Csr {
    int64_t d1 = 0, d2 = 0, nnz = 0;
    std::vector<int64_t> row_ptr;
    std::vector<int32_t> item;
    std::vector<double> rating;
};

struct Triples { std::vector<int32_t> u, i; std::vector<double> r; };

inline void parse_chunk(const char *p, const char *end, Triples &out) {
    while (p < end) {
        while (p < end && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n')) ++p;
        if (p >= end) break;
        long a = 0, b = 0;
        bool neg = false;
        if (*p == '-') { neg = true; ++p; }
        while (p < end && *p >= '0' && *p <= '9') a = a * 10 + (*p++ - '0');
        if (neg) a = -a;
        while (p < end && (*p == ' ' || *p == '\t')) ++p;
        neg = false;
        if (p < end && *p == '-') { neg = true; ++p; }
        while (p < end && *p >= '0' && *p <= '9') b = b * 10 + (*p++ - '0');
        if (neg) b = -b;
        while (p < end && (*p == ' ' || *p == '\t')) ++p;
        // rating: fast path for plain decimals, strtod for anything else (exponents, inf, ...)
        const char *q = p;
        while (q < end && *q != '\n') ++q;
        char buf[64];
        size_t len = (size_t)(q - p);
        if (len > 63) len = 63;
        memcpy(buf, p, len); buf[len] = 0;
        const double v = strtod(buf, nullptr);
        out.u.push_back((int32_t)(a - 1)); out.i.push_back((int32_t)(b - 1)); out.r.push_back(v);
        p = q;
    }
}

// parses at most `limit` entries (the reference reads exactly the count given in meta)
inline Triples parse_file(const std::string &path, int64_t limit) {
    int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) throw std::runtime_error("cannot open " + path);
    struct stat st; fstat(fd, &st);
    const size_t size = (size_t)st.st_size;
    Triples all;
    if (size == 0) { close(fd); return all; }
    const char *base = (const char *)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (base == MAP_FAILED) { close(fd); throw std::runtime_error("mmap failed for " + path); }
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    if ((size_t)nt > size / 65536 + 1) nt = (int)(size / 65536 + 1);
    std::vector<size_t> cut(nt + 1, size);
    cut[0] = 0;
    for (int t = 1; t < nt; ++t) {
        size_t c = size / nt * t;
        while (c < size && base[c] != '\n') ++c;
        cut[t] = c < size ? c + 1 : size;
    }
    std::vector<Triples> part(nt);
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < nt; ++t) {
        part[t].u.reserve((cut[t + 1] - cut[t]) / 12 + 16);
        parse_chunk(base + cut[t], base + cut[t + 1], part[t]);
    }
    munmap((void *)base, size); close(fd);
    size_t total = 0;
    for (auto &p : part) total += p.u.size();
    if (limit >= 0 && (size_t)limit < total) total = (size_t)limit;
    all.u.resize(total); all.i.resize(total); all.r.resize(total);
    size_t off = 0;
    for (auto &p : part) {
        const size_t n = std::min(p.u.size(), total - off);
        std::copy(p.u.begin(), p.u.begin() + n, all.u.begin() + off);
        std::copy(p.i.begin(), p.i.begin() + n, all.i.begin() + off);
        std::copy(p.r.begin(), p.r.begin() + n, all.r.begin() + off);
        off += n;
        if (off >= total) break;
    }
    return all;
}

// CSR sorted by (user, item): the order of smat_t::load_from_iterator util.h:240-247
inline Csr build_train_csr(int64_t d1, int64_t d2, const Triples &t) {
    Csr X; X.d1 = d1; X.d2 = d2; X.nnz = (int64_t)t.u.size();
    X.row_ptr.assign(d1 + 1, 0);
    for (size_t e = 0; e < t.u.size(); ++e) {
        if (t.u[e] < 0 || t.u[e] >= d1 || t.i[e] < 0 || t.i[e] >= d2) throw std::runtime_error("rating id out of range");
        X.row_ptr[t.u[e] + 1]++;
    }
    for (int64_t u = 0; u < d1; ++u) X.row_ptr[u + 1] += X.row_ptr[u];
    X.item.resize(X.nnz); X.rating.resize(X.nnz);
    std::vector<int64_t> fill(X.row_ptr.begin(), X.row_ptr.end() - 1);
    for (size_t e = 0; e < t.u.size(); ++e) { const int64_t p = fill[t.u[e]]++; X.item[p] = t.i[e]; X.rating[p] = t.r[e]; }
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t u = 0; u < d1; ++u) {
        const int64_t a = X.row_ptr[u], b = X.row_ptr[u + 1];
        bool sorted = true;
        for (int64_t e = a + 1; e < b; ++e) if (X.item[e] < X.item[e - 1]) { sorted = false; break; }
        if (sorted) continue;
        std::vector<std::pair<int32_t, double>> tmp((size_t)(b - a));
        for (int64_t e = a; e < b; ++e) tmp[e - a] = std::make_pair(X.item[e], X.rating[e]);
        std::stable_sort(tmp.begin(), tmp.end(), [](const std::pair<int32_t, double> &x, const std::pair<int32_t, double> &y) { return x.first < y.first; });
        for (int64_t e = a; e < b; ++e) { X.item[e] = tmp[e - a].first; X.rating[e] = tmp[e - a].second; }
    }
    return X;
}

// convert(testset_t&, d1, d2) util.cpp:250-274, literally: file order, entries lumped while T[cc].i <= j
inline Csr build_test_csr(int64_t d1, int64_t d2, const Triples &t) {
    Csr X; X.d1 = d1; X.d2 = d2;
    const int64_t nnz = (int64_t)t.u.size();
    X.row_ptr.assign(d1 + 1, 0); X.item.resize(nnz); X.rating.resize(nnz);
    int64_t cc = 0;
    for (int64_t j = 0; j < d1; ++j) {
        X.row_ptr[j] = cc;
        for (; cc < nnz; ++cc) {
            if (t.u[cc] > j) break;
            X.item[cc] = t.i[cc]; X.rating[cc] = t.r[cc];
        }
    }
    X.row_ptr[d1] = cc; X.nnz = cc;
    X.item.resize(cc); X.rating.resize(cc);
    return X;
}

struct DataDir { Csr train, test; bool has_test = false; };

inline DataDir load_dir(const std::string &dir) {
    FILE *fp = fopen((dir + "/meta").c_str(), "r");
    if (!fp) throw std::runtime_error("cannot open " + dir + "/meta");
    long m = 0, n = 0, nnz = 0, nnz_t = 0;
    char buf[1024], buf2[1024];
    if (fscanf(fp, "%ld %ld", &m, &n) != 2 || fscanf(fp, "%ld %1023s", &nnz, buf) != 2) { fclose(fp); throw std::runtime_error("bad meta file"); }
    const bool has_test = fscanf(fp, "%ld %1023s", &nnz_t, buf2) == 2;
    fclose(fp);
    DataDir d;
    d.train = build_train_csr(m, n, parse_file(dir + "/" + buf, nnz));
    d.has_test = has_test;
    if (has_test) d.test = build_test_csr(m, n, parse_file(dir + "/" + buf2, nnz_t));
    else { d.test.d1 = m; d.test.d2 = n; d.test.row_ptr.assign(m + 1, 0); }
    return d;
}

}  // namespace pcrhost
