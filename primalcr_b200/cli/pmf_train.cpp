// primalcr-train -- our own host CLI with the reference's command line (omp-pmf-train, pmf-train.cpp:8-135):
//
//   primalcr-train [options] data_dir [model_filename]
//     -s type   1 -- PrimalCR, 2 -- PrimalCR++ (default 2)       -k rank (default 10)
//     -n threads (accepted; host threads for loading only)        -l lambda (default 5000)
//     -t max_iter (default 10)                                    -p do_predict (default 1)
//
// Same data directory format (meta + ratings files), same stdout lines, same U.txt / V.txt (U<lambda>.txt for -s 1)
// side files and the same binary model layout (util.cpp:30-51) as the reference; the solver runs on the GPU(s)
// through the C ABI (PRIMALCR_GPUS=N shards users over N GPUs).  Differences on purpose: a parallel mmap loader
// instead of fgets/sscanf (host/loader.hpp), "-s 0" (CCD++) is not offered, and -w (warm start) is an extension:
//     -w model_file : start from the U, V of an existing model file instead of initial()
#include "../host/driver.hpp"
#include "../host/loader.hpp"
#include "../host/textio.hpp"

#include <cstring>
#include <fstream>
#include <string>
#ifdef _OPENMP
#include <omp.h>
#endif

static void exit_with_help() {
    printf(
        "Usage: primalcr-train [options] data_dir [model_filename]\n"
        "options:\n"
        "    -s type : set type of solver (default 2)\n"
        "    \t 1 -- PirmalCR\n"
        "    \t 2 -- PrimalCR++\n"
        "    -k rank : set the rank (default 10)\n"
        "    -n threads : set the number of host threads used for loading (default 4)\n"
        "    -l lambda : set the regularization parameter lambda (default 5000)\n"
        "    -t max_iter: set the number of iterations (default 10)\n"
        "    -p do_predict: compute training/testing error & NDCG at each iteration or not (default 1)\n"
        "    -w model : warm start from an existing model file (extension)\n"
        "environment: PRIMALCR_GPUS=N shards the users over N GPUs of this box (default 1)\n");
    exit(1);
}

static double wall() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

static void write_text(const std::string &path, const std::vector<double> &M, long rows, int k) {
    pcrhost::write_text_matrix(path.c_str(), M.data(), rows, k);
}

static void write_matrix(FILE *fp, const std::vector<double> &M, long rows, long k) {   // save_mat_t util.cpp:30-51
    fwrite(&rows, sizeof(long), 1, fp);
    fwrite(&k, sizeof(long), 1, fp);
    fwrite(M.data(), sizeof(double), (size_t)rows * k, fp);
}

static bool read_matrix(FILE *fp, std::vector<double> &M, long &rows, long &k) {         // load_mat_t util.cpp:56-79
    if (fread(&rows, sizeof(long), 1, fp) != 1 || fread(&k, sizeof(long), 1, fp) != 1) return false;
    M.resize((size_t)rows * k);
    return fread(M.data(), sizeof(double), (size_t)rows * k, fp) == (size_t)rows * k;
}

int main(int argc, char **argv) {
    primalcr_config cfg; primalcr_default_config(&cfg);
    int threads = 4;
    std::string warm;
    int i;
    for (i = 1; i < argc; i++) {
        if (argv[i][0] != '-') break;
        if (++i >= argc) exit_with_help();
        switch (argv[i - 1][1]) {
            case 's': cfg.solver = atoi(argv[i]); break;
            case 'k': cfg.k = atoi(argv[i]); break;
            case 'n': threads = atoi(argv[i]); cfg.threads = threads; break;
            case 'l': cfg.lambda = atof(argv[i]); break;
            case 't': cfg.maxiter = atoi(argv[i]); break;
            case 'p': cfg.do_predict = atoi(argv[i]); break;
            case 'w': warm = argv[i]; break;
            case 'r': case 'T': case 'e': case 'B': case 'm': case 'u': case 'd': case 'q': case 'N':
                break;      // CCD++/DSGD knobs of the reference parser (pmf-train.cpp:59-102): accepted, unused
            default:
                fprintf(stderr, "unknown option: -%c\n", argv[i - 1][1]);
                exit_with_help();
        }
    }
    if (i >= argc) exit_with_help();
    if (cfg.solver != PRIMALCR_SOLVER_PCR && cfg.solver != PRIMALCR_SOLVER_PCRPP) {
        fprintf(stderr, "Error: wrong solver type (%d)!\n", cfg.solver);
        return 0;
    }
    std::string input = argv[i], model;
    if (i < argc - 1) model = argv[i + 1];
    else {                                                   // pmf-train.cpp:122-133
        std::string p = input;
        while (!p.empty() && p.back() == '/') p.pop_back();
        const size_t s = p.rfind('/');
        model = (s == std::string::npos ? p : p.substr(s + 1)) + ".model";
    }
    // The reference opens (and truncates) the model file before loading anything (pmf-train.cpp:252-259).  We write to
    // <model>.tmp and rename() after a successful solve instead: `-w m.model data m.model` keeps training the same file,
    // and a bad data dir or a failed GPU init never leaves a zero-length file over a good model.
    const std::string model_tmp = model + ".tmp";
    FILE *model_fp = fopen(model_tmp.c_str(), "wb");
    if (!model_fp) { fprintf(stderr, "can't open output file %s\n", model.c_str()); exit(1); }
#ifdef _OPENMP
    omp_set_num_threads(threads > 0 ? threads : 1);
#endif
    pcrhost::Lap lap;                 // PRIMALCR_VERBOSE_SETUP=1: wall-time split load / init / solve / write on stderr
    pcrhost::DataDir data;
    try { data = pcrhost::load_dir(input); }
    catch (const std::exception &ex) { fprintf(stderr, "primalcr-train: %s\n", ex.what()); fclose(model_fp); remove(model_tmp.c_str()); return 1; }
    lap("cli: load data directory");
    const pcrhost::Csr &X = data.train, &T = data.test;
    const int k = cfg.k;
    std::vector<double> U((size_t)X.d1 * k), V((size_t)X.d2 * k);
    if (warm.empty()) {
        primalcr_reference_init(U.data(), X.d1, k);          // initial() util.cpp:80-93 (same libstdc++ stream)
        primalcr_reference_init(V.data(), X.d2, k);
    } else {
        FILE *wf = fopen(warm.c_str(), "rb");
        long r1 = 0, k1 = 0, r2 = 0, k2 = 0;
        if (!wf || !read_matrix(wf, U, r1, k1) || !read_matrix(wf, V, r2, k2) || r1 != X.d1 || r2 != X.d2 || k1 != k || k2 != k) {
            fprintf(stderr, "can't warm start from %s (need %ld x %d and %ld x %d)\n", warm.c_str(), (long)X.d1, k, (long)X.d2, k);
            fclose(model_fp); remove(model_tmp.c_str());
            exit(1);
        }
        fclose(wf);
    }
    lap("cli: initial U, V");
    std::cout << "the rank is " << k << std::endl;
    std::cout << "the number of rows is " << X.d1 << " and the number of cols is " << X.d2 << std::endl;
    if (cfg.solver == PRIMALCR_SOLVER_PCR) std::cout << "nnz: " << X.nnz << std::endl;      // pmf-train.cpp:201
    else { std::cout << X.nnz << std::endl; std::cout << "starts!" << std::endl; }          // pmf-train.cpp:270-271
    const double t0 = wall();
    pcrhost::FlatCsr fx{X.d1, X.d2, X.nnz, X.row_ptr.data(), X.item.data(), X.rating.data()};
    pcrhost::FlatCsr ft{T.d1, T.d2, T.nnz, T.row_ptr.data(), T.item.data(), T.rating.data()};
    pcrhost::solve(cfg, fx, ft, U.data(), V.data(), pcrhost::gpus_from_env());
    printf("Wall-time: %lg secs\n", wall() - t0);
    lap("cli: solve (see host laps)");
    const bool pp = cfg.solver == PRIMALCR_SOLVER_PCRPP;
    const std::string suffix = pp ? "" : std::to_string((int)cfg.lambda);
    std::cout << "U matrix of size " << X.d1 << ", " << k << std::endl;
    if (!getenv("PRIMALCR_NO_TEXT_DUMP")) write_text("U" + suffix + ".txt", U, X.d1, k);
    std::cout << "V matrix of size " << X.d2 << ", " << k << std::endl;
    if (!getenv("PRIMALCR_NO_TEXT_DUMP")) write_text("V" + suffix + ".txt", V, X.d2, k);
    lap("cli: U.txt / V.txt text dumps");
    write_matrix(model_fp, U, X.d1, k);
    write_matrix(model_fp, V, X.d2, k);
    if (fclose(model_fp) != 0 || rename(model_tmp.c_str(), model.c_str()) != 0) {
        fprintf(stderr, "can't write model file %s\n", model.c_str());
        return 1;
    }
    lap("cli: model file");
    return 0;
}
