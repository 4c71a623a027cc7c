// libmlg_ingest.so: reads file (FASTQ / single-line FASTA, optionally gzip) -> packed batches for
// mlg_query_push_packed_nruns().  Declared in include/metalign_b200_ingest.h; replaces the input side of
// `kmc -k60 -fq|-fa` (scripts/select_db.py:46-52 of the reference).  Host code only.
//
//   reader thread    gzread() (transparent for plain files) into blocks
//   scanner thread   cuts blocks into lines, keeps the sequence lines as (offset, length) pairs
//   mlgi_next()      takes sequence lines until the batch is full, prefix-sums their lengths into read offsets and
//                    lets `threads` workers pack disjoint ranges of the output stream
#include <zlib.h>
#include <fcntl.h>
#include <unistd.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <errno.h>
#include <algorithm>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "../../include/metalign_b200_ingest.h"

#define MLGI_API __attribute__((visibility("default")))

namespace {

thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

template <typename T>
class BoundedQueue {
public:
    explicit BoundedQueue(size_t cap) : cap_(cap) {}
    // false when the queue was closed for good by the consumer
    bool push(T v) {
        std::unique_lock<std::mutex> g(mu_);
        not_full_.wait(g, [&] { return q_.size() < cap_ || abandoned_; });
        if (abandoned_) return false;
        q_.push_back(std::move(v));
        not_empty_.notify_one();
        return true;
    }
    // false when the producer has finished and nothing is left
    bool pop(T& out) {
        std::unique_lock<std::mutex> g(mu_);
        not_empty_.wait(g, [&] { return !q_.empty() || done_; });
        if (q_.empty()) return false;
        out = std::move(q_.front());
        q_.pop_front();
        not_full_.notify_one();
        return true;
    }
    void finish() { std::lock_guard<std::mutex> g(mu_); done_ = true; not_empty_.notify_all(); }
    void abandon() { std::lock_guard<std::mutex> g(mu_); abandoned_ = true; done_ = true; not_full_.notify_all(); not_empty_.notify_all(); }
private:
    std::mutex mu_;
    std::condition_variable not_full_, not_empty_;
    std::deque<T> q_;
    size_t cap_;
    bool done_ = false, abandoned_ = false;
};

// a block of file text with HEAD spare bytes in front of it, so that the unterminated tail of the previous block can
// be put in front without copying the block (uninitialised storage: no zero-fill pass over every block)
constexpr size_t HEAD = 1u << 16;
struct TextBuf {
    std::unique_ptr<char[]> mem;
    size_t off = HEAD, n = 0;                 // text = mem[off, off + n)
    const char* data() const { return mem.get() + off; }
    size_t size() const { return n; }
};
struct RawBlock { std::shared_ptr<TextBuf> buf; };
struct Lines {
    std::shared_ptr<TextBuf> text;
    std::vector<uint32_t> start, len;       // sequence lines inside text
};

unsigned char g_code[256];
struct CodeInit {
    CodeInit() {
        memset(g_code, 4, sizeof(g_code));
        const char* s = "ACGT";
        for (int i = 0; i < 4; ++i) { g_code[(unsigned char)s[i]] = (unsigned char)i; g_code[(unsigned char)(s[i] + 32)] = (unsigned char)i; }
    }
} g_code_init;

}  // namespace

struct mlgi_reader {
    gzFile fh = nullptr;
    int fd = -1;                        // plain (not gzip) files are read with read(2): no inflate layer in the way
    int type = MLGI_FASTQ;
    int threads = 1;
    size_t block_bytes = 8u << 20;
    BoundedQueue<RawBlock> q_raw{4};
    BoundedQueue<Lines> q_lines{4};
    std::thread t_read, t_scan;
    std::string io_error;               // set by the reader thread before it finishes the queue
    std::mutex err_mu;
    // consumer state
    Lines cur;
    size_t cur_i = 0;
    bool eof = false;
    uint64_t tot_reads = 0, tot_bases = 0, tot_text = 0;

    void read_loop() {
        for (;;) {
            RawBlock b;
            b.buf = std::make_shared<TextBuf>();
            b.buf->mem.reset(new char[HEAD + block_bytes]);
            char* dst = b.buf->mem.get() + HEAD;
            size_t got = 0;
            while (got < block_bytes) {
                int want = (int)std::min<size_t>(block_bytes - got, 1u << 30);
                long n = fd >= 0 ? (long)read(fd, dst + got, (size_t)want) : (long)gzread(fh, dst + got, (unsigned)want);
                if (n < 0) {
                    int errnum = 0;
                    const char* m = fd >= 0 ? strerror(errno) : gzerror(fh, &errnum);
                    std::lock_guard<std::mutex> g(err_mu);
                    io_error = m ? m : "read failed";
                    q_raw.finish();
                    return;
                }
                if (n == 0) break;
                got += (size_t)n;
            }
            b.buf->n = got;
            const bool last = got < block_bytes;
            if (got && !q_raw.push(std::move(b))) return;
            if (last) { q_raw.finish(); return; }
        }
    }

    void scan_loop() {
        std::string carry;               // the unterminated tail of the previous block
        uint64_t line_no = 0;            // lines completed so far (FASTQ: record = 4 lines)
        RawBlock b;
        auto emit_lines = [&](std::shared_ptr<TextBuf>& text, bool final_block, Lines& out) {
            const char* base = text->data();
            const size_t n = text->size();
            size_t p = 0;
            while (p < n) {
                const char* nl = (const char*)memchr(base + p, '\n', n - p);
                size_t e;
                if (nl) e = (size_t)(nl - base);
                else if (final_block) e = n;
                else break;                                   // incomplete line: carried into the next block
                size_t le = e;
                if (le > p && base[le - 1] == '\r') --le;
                bool is_seq;
                if (type == MLGI_FASTQ) is_seq = (line_no & 3u) == 1u;
                else is_seq = le > p && base[p] != '>' && base[p] != ';';
                if (is_seq) { out.start.push_back((uint32_t)p); out.len.push_back((uint32_t)(le - p)); }
                ++line_no;
                p = nl ? e + 1 : n;
            }
            return p;                                         // bytes consumed
        };
        bool more = true;
        while (more) {
            more = q_raw.pop(b);
            std::shared_ptr<TextBuf> text;
            if (more && carry.size() <= HEAD) {                // the usual case: the carried tail fits in front of the block
                text = b.buf;
                memcpy(text->mem.get() + HEAD - carry.size(), carry.data(), carry.size());
                text->off = HEAD - carry.size(); text->n += carry.size();
            } else {
                text = std::make_shared<TextBuf>();
                const size_t bn = more ? b.buf->n : 0;
                text->mem.reset(new char[carry.size() + bn + 1]);
                text->off = 0; text->n = carry.size() + bn;
                if (!carry.empty()) memcpy(text->mem.get(), carry.data(), carry.size());
                if (bn) memcpy(text->mem.get() + carry.size(), b.buf->data(), bn);   // (or, at the end, just the last unterminated line)
            }
            if (text->size() >= 0xFFFFFFF0ull) { std::lock_guard<std::mutex> g(err_mu); io_error = "a single line exceeds 4 GiB"; break; }
            Lines out;
            out.text = text;
            const size_t used = emit_lines(text, !more, out);
            carry.assign(text->data() + used, text->size() - used);
            if (!out.start.empty() && !q_lines.push(std::move(out))) return;
        }
        q_lines.finish();
    }
};

namespace {

// pack reads [a, b) of the batch; the worker owns stream bases [off[a], off[b])
struct PackJob {
    const char* const* ptr; const uint32_t* len; const uint64_t* off;
    size_t a, b;
    uint8_t* bases;
    std::vector<uint32_t> runs;      // (start, length) pairs found in this range
};

void pack_range(PackJob& j) {
    if (j.a >= j.b) return;
    uint64_t pos = j.off[j.a];
    uint8_t* out = j.bases + (pos >> 2);
    unsigned fill = (unsigned)(pos & 3u);      // bases already present in the current output byte (owned by a neighbour)
    unsigned acc = 0;
    bool first_partial = fill != 0;
    uint64_t run_start = 0; uint32_t run_len = 0;
    for (size_t i = j.a; i < j.b; ++i) {
        const unsigned char* s = (const unsigned char*)j.ptr[i];
        const uint32_t L = j.len[i];
        for (uint32_t k = 0; k < L; ++k) {
            unsigned c = g_code[s[k]];
            if (c > 3u) {
                if (run_len && run_start + run_len == pos) ++run_len;
                else { if (run_len) { j.runs.push_back((uint32_t)run_start); j.runs.push_back(run_len); } run_start = pos; run_len = 1; }
                c = 0;
            }
            acc = (acc << 2) | c;
            ++pos;
            if (++fill == 4) {
                if (first_partial) { __atomic_fetch_or(out, (uint8_t)acc, __ATOMIC_RELAXED); first_partial = false; }
                else *out = (uint8_t)acc;
                ++out; fill = 0; acc = 0;
            }
        }
    }
    if (run_len) { j.runs.push_back((uint32_t)run_start); j.runs.push_back(run_len); }
    if (fill) {     // trailing partial byte: shared with the next worker (or zero padding)
        const uint8_t v = (uint8_t)(acc << (2 * (4 - fill)));
        __atomic_fetch_or(out, v, __ATOMIC_RELAXED);
    }
}

}  // namespace

extern "C" {

MLGI_API const char* mlgi_last_error(void) { return g_err; }

MLGI_API int mlgi_open(const char* path, int input_type, int threads, mlgi_reader** out) {
    if (!path || !out) { set_error("null argument"); return -2; }
    if (input_type != MLGI_FASTQ && input_type != MLGI_FASTA) { set_error("input_type must be MLGI_FASTQ or MLGI_FASTA"); return -2; }
    int fd = open(path, O_RDONLY);
    if (fd < 0) { set_error("cannot open %s", path); return -3; }
    unsigned char magic[2] = {0, 0};
    const bool gz = pread(fd, magic, 2, 0) == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
    gzFile fh = nullptr;
    if (gz) {
        fh = gzdopen(fd, "rb");
        if (!fh) { close(fd); set_error("cannot open %s", path); return -3; }
        gzbuffer(fh, 1u << 20);
        fd = -1;
    }
    mlgi_reader* r = new mlgi_reader();
    r->fh = fh; r->fd = fd; r->type = input_type;
    int hw = (int)std::thread::hardware_concurrency();
    if (hw < 1) hw = 1;
    r->threads = threads > 0 ? threads : hw;
    if (const char* s = getenv("MLGI_BLOCK_BYTES")) { long v = atol(s); if (v >= 16) r->block_bytes = (size_t)v; }
    r->t_read = std::thread([r] { r->read_loop(); });
    r->t_scan = std::thread([r] { r->scan_loop(); });
    *out = r;
    return 0;
}

MLGI_API int mlgi_next(mlgi_reader* r, uint8_t* bases, uint64_t cap_bases_bytes, uint32_t* nruns, uint64_t cap_runs, uint64_t* off,
                       uint64_t max_reads, uint64_t max_bases, uint64_t* n_reads, uint64_t* n_runs) {
    if (!r || !bases || !off || !n_reads || !n_runs) { set_error("null argument"); return -2; }
    if (max_reads == 0 || max_bases == 0 || max_bases >= 0xFFFFFFFFull) { set_error("max_reads must be > 0 and 0 < max_bases < 2^32"); return -2; }
    if (cap_bases_bytes < max_bases / 4 + 32) { set_error("bases buffer too small for max_bases"); return -2; }
    *n_reads = 0; *n_runs = 0;
    std::vector<const char*> ptr;
    std::vector<uint32_t> len;
    std::vector<std::shared_ptr<TextBuf>> keep;
    uint64_t nb = 0;
    while (ptr.size() < max_reads) {
        if (r->cur_i >= r->cur.start.size()) {
            if (r->eof) break;
            Lines nx;
            if (!r->q_lines.pop(nx)) { r->eof = true; break; }
            r->tot_text += nx.text->size();
            r->cur = std::move(nx); r->cur_i = 0;
        }
        if (keep.empty() || keep.back() != r->cur.text) keep.push_back(r->cur.text);
        bool full = false;
        while (r->cur_i < r->cur.start.size() && ptr.size() < max_reads) {
            const uint32_t L = r->cur.len[r->cur_i];
            if (L > max_bases) { set_error("a read of %u bases does not fit a batch of %llu bases", L, (unsigned long long)max_bases); return -2; }
            if (nb + L > max_bases) { full = true; break; }
            ptr.push_back(r->cur.text->data() + r->cur.start[r->cur_i]);
            len.push_back(L);
            nb += L;
            ++r->cur_i;
        }
        if (full) break;
    }
    {
        std::lock_guard<std::mutex> g(r->err_mu);
        if (!r->io_error.empty()) { set_error("read error: %s", r->io_error.c_str()); return -3; }
    }
    const size_t n = ptr.size();
    if (n == 0) return 0;
    off[0] = 0;
    for (size_t i = 0; i < n; ++i) off[i + 1] = off[i] + len[i];
    // worker ranges: equal shares of the bases
    const int T = (int)std::min<size_t>((size_t)r->threads, std::max<size_t>(1, nb / 65536));
    std::vector<PackJob> jobs((size_t)T);
    size_t a = 0;
    for (int t = 0; t < T; ++t) {
        const uint64_t target = nb * (uint64_t)(t + 1) / (uint64_t)T;
        size_t b = (t == T - 1) ? n : (size_t)(std::upper_bound(off, off + n + 1, target) - off) - 1;
        if (b < a) b = a;
        jobs[(size_t)t] = PackJob{ptr.data(), len.data(), off, a, b, bases, {}};
        a = b;
    }
    // bytes two workers may both touch (and the padding) start out zero
    for (int t = 0; t < T; ++t) {
        const PackJob& j = jobs[(size_t)t];
        bases[off[j.a] >> 2] = 0;
        bases[off[j.b] >> 2] = 0;
    }
    const uint64_t used = (nb + 3) / 4, padded = (used + 15) / 16 * 16 + 16;
    memset(bases + used, 0, (size_t)std::min<uint64_t>(padded, cap_bases_bytes) - used);
    if (T == 1) pack_range(jobs[0]);
    else {
        std::vector<std::thread> th;
        for (int t = 1; t < T; ++t) th.emplace_back([&jobs, t] { pack_range(jobs[(size_t)t]); });
        pack_range(jobs[0]);
        for (auto& x : th) x.join();
    }
    // N runs of the workers, in stream order; runs that touch across a worker boundary are merged
    uint64_t nr = 0;
    for (int t = 0; t < T; ++t) {
        const std::vector<uint32_t>& v = jobs[(size_t)t].runs;
        for (size_t i = 0; i + 1 < v.size(); i += 2) {
            if (nr && (uint64_t)nruns[2 * (nr - 1)] + nruns[2 * (nr - 1) + 1] == v[i]) { nruns[2 * (nr - 1) + 1] += v[i + 1]; continue; }
            if (nr >= cap_runs || !nruns) { set_error("more than %llu N runs in one batch", (unsigned long long)cap_runs); return -2; }
            nruns[2 * nr] = v[i]; nruns[2 * nr + 1] = v[i + 1];
            ++nr;
        }
    }
    *n_reads = n; *n_runs = nr;
    r->tot_reads += n; r->tot_bases += nb;
    return 1;
}

MLGI_API int mlgi_stats(mlgi_reader* r, uint64_t* reads, uint64_t* bases, uint64_t* text_bytes) {
    if (!r) { set_error("null reader"); return -2; }
    if (reads) *reads = r->tot_reads;
    if (bases) *bases = r->tot_bases;
    if (text_bytes) *text_bytes = r->tot_text;
    return 0;
}

MLGI_API void mlgi_close(mlgi_reader* r) {
    if (!r) return;
    r->q_lines.abandon();
    r->q_raw.abandon();
    if (r->t_scan.joinable()) r->t_scan.join();
    if (r->t_read.joinable()) r->t_read.join();
    if (r->fh) gzclose(r->fh);
    if (r->fd >= 0) close(r->fd);
    delete r;
}

}  // extern "C"
