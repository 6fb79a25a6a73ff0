// ingest.cu — Wannier90 text files straight into the C-ABI memory images (host code).
//
// The reference delegates parsing to WannierIO.jl (third party, not in the tree: `read_mmn`, `read_amn`,
// `read_eig`, called from src/io/w90/model.jl:16-94).  At the north-star shape the .mmn is ~13 GB of text
// (49 152 pairs x 16 384 lines), so the parser is native and parallel: the file is mapped, split into byte
// ranges, every thread counts the newlines of its range (pass 1), a prefix sum turns that into the global
// line number at which each range starts, and every thread then parses its own lines directly into the
// final layout (pass 2) — line number -> (pair, element) is arithmetic because every pair is one header
// line followed by nb*nb data lines.  Numbers go through std::from_chars (correctly rounded, like Julia's
// `parse(Float64, .)`), with a strtod fallback for Fortran 'D' exponents.
//
// Formats (restated from the files; tests/golden/w90_* hold the reference's own fixtures):
//   .mmn  comment; `nb nk nn`; per pair `ik ikpb G1 G2 G3` (1-based) then nb*nb lines `re im`, m fastest; the
//         order in which the pairs of a k-point appear defines b (src/kpoints/kstencil.jl:68-85)
//   .amn  comment; `nb nk nw`; lines `m n ik re im`
//   .eig  lines `m ik eps`
#include "../../include/wannier_b200.h"

#include <algorithm>
#include <atomic>
#include <charconv>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

thread_local std::string g_io_error;

struct Mapped {
    const char* p = nullptr;
    size_t n = 0;
    int fd = -1;
    bool open(const char* path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) { g_io_error = std::string("cannot open ") + path; return false; }
        struct stat st;
        if (fstat(fd, &st) != 0 || st.st_size <= 0) { g_io_error = std::string("cannot stat ") + path; return false; }
        n = (size_t)st.st_size;
        void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { g_io_error = std::string("cannot map ") + path; return false; }
        madvise(m, n, MADV_SEQUENTIAL);
        p = (const char*)m;
        return true;
    }
    ~Mapped() {
        if (p) munmap((void*)p, n);
        if (fd >= 0) ::close(fd);
    }
};

inline const char* skip_ws(const char* s, const char* e) {
    while (s < e && (*s == ' ' || *s == '\t' || *s == '\r')) s++;
    return s;
}
inline const char* line_end(const char* s, const char* e) {
    const char* q = (const char*)memchr(s, '\n', (size_t)(e - s));
    return q ? q : e;
}
inline bool parse_int(const char*& s, const char* e, long& v) {
    s = skip_ws(s, e);
    if (s < e && *s == '+') s++;
    auto r = std::from_chars(s, e, v);
    if (r.ec != std::errc()) return false;
    s = r.ptr;
    return true;
}
inline bool parse_double(const char*& s, const char* e, double& v) {
    s = skip_ws(s, e);
    if (s < e && *s == '+') s++;
    auto r = std::from_chars(s, e, v);
    if (r.ec != std::errc()) return false;
    if (r.ptr < e && (*r.ptr == 'D' || *r.ptr == 'd')) {   // Fortran double-precision exponent
        char buf[64];
        size_t len = 0;
        const char* q = s;
        while (q < e && len < 63 && *q != ' ' && *q != '\t' && *q != '\n' && *q != '\r') {
            buf[len++] = (*q == 'D' || *q == 'd') ? 'e' : *q;
            q++;
        }
        buf[len] = 0;
        char* endp = nullptr;
        v = strtod(buf, &endp);
        if (endp == buf) return false;
        s = q;
        return true;
    }
    s = r.ptr;
    return true;
}

int default_threads() {
    unsigned h = std::thread::hardware_concurrency();
    if (h == 0) h = 4;
    return (int)std::min(h, 32u);
}

// [begin, end) byte ranges aligned to line starts
std::vector<size_t> split_lines(const char* p, size_t lo, size_t hi, int parts) {
    std::vector<size_t> b(parts + 1);
    b[0] = lo; b[parts] = hi;
    for (int i = 1; i < parts; i++) {
        size_t pos = lo + (hi - lo) / parts * i;
        if (pos < b[i - 1]) pos = b[i - 1];
        const char* q = (const char*)memchr(p + pos, '\n', hi - pos);
        b[i] = q ? (size_t)(q - p) + 1 : hi;
    }
    return b;
}

template <class F>
void run_threads(int n, F&& f) {
    std::vector<std::thread> th;
    for (int t = 1; t < n; t++) th.emplace_back([&f, t] { f(t); });
    f(0);
    for (auto& x : th) x.join();
}

// offset of the first byte after `count` lines starting at `lo`
size_t after_lines(const char* p, size_t lo, size_t hi, int count) {
    size_t pos = lo;
    for (int i = 0; i < count && pos < hi; i++) {
        const char* q = (const char*)memchr(p + pos, '\n', hi - pos);
        pos = q ? (size_t)(q - p) + 1 : hi;
    }
    return pos;
}

int read_header3(const Mapped& f, long out[3], size_t* body) {
    const size_t l2 = after_lines(f.p, 0, f.n, 1);
    const char* s = f.p + l2;
    const char* e = line_end(s, f.p + f.n);
    for (int i = 0; i < 3; i++)
        if (!parse_int(s, e, out[i]) || out[i] <= 0) { g_io_error = "bad size line (expected three positive integers)"; return WB200_ERR_ARG; }
    *body = after_lines(f.p, l2, f.n, 1);
    return WB200_OK;
}

}  // namespace

extern "C" const char* wb200_io_last_error(void) { return g_io_error.c_str(); }

extern "C" int wb200_mmn_header(const char* path, int* nb, int* nk, int* nn) {
    if (!path || !nb || !nk || !nn) return WB200_ERR_ARG;
    Mapped f;
    if (!f.open(path)) return WB200_ERR_ARG;
    long h[3]; size_t body;
    if (int s = read_header3(f, h, &body)) return s;
    *nb = (int)h[0]; *nk = (int)h[1]; *nn = (int)h[2];
    return WB200_OK;
}

extern "C" int wb200_mmn_read(const char* path, int nb, int nk, int nn, double* M, int32_t* kpb_k, int32_t* kpb_G,
                              int nthreads) {
    if (!path || !M || !kpb_k || !kpb_G || nb <= 0 || nk <= 0 || nn <= 0) return WB200_ERR_ARG;
    Mapped f;
    if (!f.open(path)) return WB200_ERR_ARG;
    long h[3]; size_t body;
    if (int s = read_header3(f, h, &body)) return s;
    if (h[0] != nb || h[1] != nk || h[2] != nn) { g_io_error = "mmn: sizes differ from the header"; return WB200_ERR_ARG; }
    const long long per = 1 + (long long)nb * nb;          // lines per pair
    const long long pairs = (long long)nk * nn, need = pairs * per;
    const int T = nthreads > 0 ? nthreads : default_threads();
    const std::vector<size_t> rng = split_lines(f.p, body, f.n, T);
    // pass 1: non-empty lines per range
    std::vector<long long> cnt(T, 0);
    run_threads(T, [&](int t) {
        long long c = 0;
        size_t pos = rng[t];
        while (pos < rng[t + 1]) {
            const char* e = line_end(f.p + pos, f.p + rng[t + 1]);
            if (skip_ws(f.p + pos, e) < e) c++;
            pos = (size_t)(e - f.p) + 1;
        }
        cnt[t] = c;
    });
    std::vector<long long> first(T + 1, 0);
    for (int t = 0; t < T; t++) first[t + 1] = first[t] + cnt[t];
    if (first[T] < need) { g_io_error = "mmn: file ends before all pairs were read"; return WB200_ERR_ARG; }
    // pass 2: parse; data of pair p (file order) goes to M[p], headers to hd[p]
    std::vector<int32_t> hd((size_t)pairs * 5);
    std::atomic<int> bad{0};
    run_threads(T, [&](int t) {
        long long line = first[t];
        size_t pos = rng[t];
        while (pos < rng[t + 1] && line < need) {
            const char* s = f.p + pos;
            const char* e = line_end(s, f.p + rng[t + 1]);
            pos = (size_t)(e - f.p) + 1;
            if (skip_ws(s, e) == e) continue;
            const long long p = line / per, off = line % per;
            if (off == 0) {
                long v[5];
                for (int i = 0; i < 5; i++)
                    if (!parse_int(s, e, v[i])) { bad = 1; return; }
                for (int i = 0; i < 5; i++) hd[(size_t)p * 5 + i] = (int32_t)v[i];
            } else {
                double re, im;
                if (!parse_double(s, e, re) || !parse_double(s, e, im)) { bad = 1; return; }
                double* dst = M + 2 * ((size_t)p * nb * nb + (size_t)(off - 1));
                dst[0] = re; dst[1] = im;
            }
            line++;
        }
    });
    if (bad) { g_io_error = "mmn: malformed line"; return WB200_ERR_ARG; }
    // b index = order of appearance among the pairs of a k-point
    std::vector<int> seen(nk, 0);
    std::vector<long long> dest(pairs);
    bool canonical = true;
    for (long long p = 0; p < pairs; p++) {
        const int ik = hd[(size_t)p * 5] - 1, ikb = hd[(size_t)p * 5 + 1] - 1;
        if (ik < 0 || ik >= nk || ikb < 0 || ikb >= nk) { g_io_error = "mmn: k-point index out of range"; return WB200_ERR_ARG; }
        const int ib = seen[ik]++;
        if (ib >= nn) { g_io_error = "mmn: every k-point must have the same number of neighbours"; return WB200_ERR_ARG; }
        dest[p] = (long long)ik * nn + ib;
        canonical = canonical && dest[p] == p;
        kpb_k[dest[p]] = ikb;
        for (int a = 0; a < 3; a++) kpb_G[dest[p] * 3 + a] = hd[(size_t)p * 5 + 2 + a];
    }
    if (!canonical) {   // pairs not grouped by k-point: permute the blocks
        const size_t blk = (size_t)nb * nb * 2;
        std::vector<double> tmp((size_t)pairs * blk);
        memcpy(tmp.data(), M, sizeof(double) * tmp.size());
        for (long long p = 0; p < pairs; p++) memcpy(M + (size_t)dest[p] * blk, tmp.data() + (size_t)p * blk, sizeof(double) * blk);
    }
    return WB200_OK;
}

extern "C" int wb200_amn_header(const char* path, int* nb, int* nk, int* nw) {
    if (!path || !nb || !nk || !nw) return WB200_ERR_ARG;
    Mapped f;
    if (!f.open(path)) return WB200_ERR_ARG;
    long h[3]; size_t body;
    if (int s = read_header3(f, h, &body)) return s;
    *nb = (int)h[0]; *nk = (int)h[1]; *nw = (int)h[2];
    return WB200_OK;
}

// shared by .amn (`m n ik re im`) and .eig (`m ik eps`): every line addresses its own element
template <class F>
static int parse_lines(const Mapped& f, size_t body, int nthreads, F&& each) {
    const int T = nthreads > 0 ? nthreads : default_threads();
    const std::vector<size_t> rng = split_lines(f.p, body, f.n, T);
    std::atomic<int> bad{0};
    run_threads(T, [&](int t) {
        size_t pos = rng[t];
        while (pos < rng[t + 1]) {
            const char* s = f.p + pos;
            const char* e = line_end(s, f.p + rng[t + 1]);
            pos = (size_t)(e - f.p) + 1;
            if (skip_ws(s, e) == e) continue;
            if (!each(s, e)) { bad = 1; return; }
        }
    });
    return bad ? WB200_ERR_ARG : WB200_OK;
}

extern "C" int wb200_amn_read(const char* path, int nb, int nk, int nw, double* A) {
    if (!path || !A || nb <= 0 || nk <= 0 || nw <= 0) return WB200_ERR_ARG;
    Mapped f;
    if (!f.open(path)) return WB200_ERR_ARG;
    long h[3]; size_t body;
    if (int s = read_header3(f, h, &body)) return s;
    if (h[0] != nb || h[1] != nk || h[2] != nw) { g_io_error = "amn: sizes differ from the header"; return WB200_ERR_ARG; }
    memset(A, 0, sizeof(double) * 2 * (size_t)nk * nb * nw);
    std::atomic<long long> count{0};
    const int s = parse_lines(f, body, 0, [&](const char* p, const char* e) {
        long m, n, ik; double re, im;
        if (!parse_int(p, e, m) || !parse_int(p, e, n) || !parse_int(p, e, ik) || !parse_double(p, e, re) ||
            !parse_double(p, e, im))
            return false;
        if (m < 1 || m > nb || n < 1 || n > nw || ik < 1 || ik > nk) return false;
        double* dst = A + 2 * (((size_t)(ik - 1) * nw + (size_t)(n - 1)) * nb + (size_t)(m - 1));   // [k][n][m]
        dst[0] = re; dst[1] = im;
        count++;
        return true;
    });
    if (s != WB200_OK) { g_io_error = "amn: malformed line or index out of range"; return s; }
    if (count != (long long)nk * nb * nw) { g_io_error = "amn: wrong number of entries"; return WB200_ERR_ARG; }
    return WB200_OK;
}

extern "C" int wb200_eig_read(const char* path, int nb, int nk, double* E) {
    if (!path || !E || nb <= 0 || nk <= 0) return WB200_ERR_ARG;
    Mapped f;
    if (!f.open(path)) return WB200_ERR_ARG;
    std::atomic<long long> count{0};
    const int s = parse_lines(f, 0, 0, [&](const char* p, const char* e) {
        long m, ik; double eps;
        if (!parse_int(p, e, m) || !parse_int(p, e, ik) || !parse_double(p, e, eps)) return false;
        if (m < 1 || m > nb || ik < 1 || ik > nk) return false;
        E[(size_t)(ik - 1) * nb + (size_t)(m - 1)] = eps;
        count++;
        return true;
    });
    if (s != WB200_OK) { g_io_error = "eig: malformed line or index out of range"; return s; }
    if (count != (long long)nk * nb) { g_io_error = "eig: wrong number of entries"; return WB200_ERR_ARG; }
    return WB200_OK;
}
