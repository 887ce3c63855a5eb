// textio.hpp -- the U.txt / V.txt side files of the reference CLI (pmf-train.cpp:276-295), written in parallel.
#pragma once
#include <algorithm>
#include <cstdio>
#include <string>
#include <vector>

namespace pcrhost {

// `myfile << M[i][j]` per entry (ostream default: 6 significant digits == printf "%g"), entries separated by one space,
// one row per line.  The reference streams one number at a time (minutes for the 3.2 GB of text of the power-law shape,
// SURVEY 8f row 3); here blocks of rows are formatted on all host threads and written in order -- byte-identical output.
// Returns false when the file cannot be opened.
inline bool write_text_matrix(const char *path, const double *M, long rows, int k) {
    FILE *f = fopen(path, "w");
    if (!f) return false;
    const long block = 4096;
    const long nblocks = (rows + block - 1) / block;
    const long wave = 256;                                     // blocks formatted per parallel wave (bounds the memory held)
    std::vector<std::string> buf((size_t)std::min(wave, std::max(nblocks, 1L)));
    for (long b0 = 0; b0 < nblocks; b0 += wave) {
        const long nb = std::min(wave, nblocks - b0);
#pragma omp parallel for schedule(dynamic, 1)
        for (long b = 0; b < nb; ++b) {
            std::string &line = buf[(size_t)b];
            line.clear();
            char tmp[64];
            const long r1 = std::min(rows, (b0 + b + 1) * block);
            for (long i = (b0 + b) * block; i < r1; ++i)
                for (int j = 0; j < k; ++j) {
                    const int n = snprintf(tmp, sizeof(tmp), "%g", M[(size_t)i * k + j]);
                    line.append(tmp, (size_t)n);
                    line += (j < k - 1) ? ' ' : '\n';
                }
        }
        for (long b = 0; b < nb; ++b) fwrite(buf[(size_t)b].data(), 1, buf[(size_t)b].size(), f);
    }
    fclose(f);
    return true;
}

}  // namespace pcrhost
