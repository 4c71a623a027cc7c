// libmlg_ingest.so: reads file (FASTQ / single-line FASTA, optionally gzip) -> packed batches for
// mlg_query_push_packed_nruns().  Declared in include/metalign_b200_ingest.h; replaces the input side of
// `kmc -k60 -fq|-fa` (scripts/select_db.py:46-52 of the reference).  Host code only.
//
//   plain file       mapped (mmap): the scanner and the packers read the page cache directly, nothing is copied
//   gzip file        reader thread: the mapped file through fast_inflate.h (an in-house DEFLATE decoder: 0.45 GB/s of
//                    FASTQ text against zlib's 0.30; the CRC-32 of every member is checked by the scanner thread;
//                    MLGI_ZLIB=1 goes back to gzread) into recycled blocks -- the stream is sequential, one thread;
//                    a BGZF file (bgzip / htslib: independent members with their sizes in the header) is mapped and
//                    its members are inflated by `threads` threads at a time, CRC-checked (1.2 GB/s of text on 8 cores)
//   scanner          cuts the text into lines (AVX2 compare + movemask where the CPU has it, memchr otherwise), keeps the
//                    sequence lines as (offset, length) pairs; for a mapped file `threads` threads do it on views of
//                    `threads` blocks in two passes (count the newlines of every part, prefix sum = the line numbers a
//                    part starts with, then cut), one thread for the gzip stream
//   mlgi_next()      takes sequence lines until the batch is full, prefix-sums their lengths into read offsets and
//                    lets `threads` workers pack disjoint ranges of the output stream: 32 bases at a time with AVX2
//                    (case-folded compare for A/C/G/T, 2-bit codes from bits 1-2 of the ASCII byte, maddubs / madd /
//                    packus to 8 output bytes), 8 at a time with 64-bit SWAR otherwise, one at a time around any
//                    other symbol (N runs) and at read boundaries that are not byte-aligned in the output
// MLGI_PROFILE=1 prints where the consumer and the scanner waited; MLGI_NO_MMAP / MLGI_NO_AVX2 force the other paths.
#include <zlib.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <fcntl.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/stat.h>
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
#include <chrono>
#include "../../include/metalign_b200_ingest.h"
#include "fast_inflate.h"
#include "kmcdb.h"

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
// blocks are recycled: a fresh 8 MiB allocation per block costs a page fault (and a zero-fill by the kernel) per 4 KiB
// inside the reader's read(2), which otherwise is the fastest stage
struct BufPool {
    std::mutex mu;
    std::vector<std::unique_ptr<char[]>> free;
    size_t bytes = 0;                          // size of the pooled blocks
    std::unique_ptr<char[]> get() {
        {
            std::lock_guard<std::mutex> g(mu);
            if (!free.empty()) { auto p = std::move(free.back()); free.pop_back(); return p; }
        }
        return std::unique_ptr<char[]>(new char[bytes]);
    }
    void put(std::unique_ptr<char[]> p) {
        std::lock_guard<std::mutex> g(mu);
        if (free.size() < 16) free.push_back(std::move(p));
    }
};
// a plain file is not read at all: it is mapped, and blocks are views of the mapping (no copy of the text; the page cache
// is what the scanner and the packers read)
struct Mapping {
    void* p = nullptr; size_t n = 0;
    ~Mapping() { if (p) munmap(p, n); }
};
struct TextBuf {
    std::unique_ptr<char[]> mem;
    std::shared_ptr<BufPool> pool;            // set when mem came from the pool
    std::shared_ptr<Mapping> map;             // set when the text is a view of a mapped file ...
    const char* ext = nullptr;                // ... starting here
    size_t off = HEAD, n = 0;                 // text = mem[off, off + n)
    const char* data() const { return ext ? ext : mem.get() + off; }
    size_t size() const { return n; }
    ~TextBuf() { if (pool && mem) pool->put(std::move(mem)); }
};
// gzip members that end inside a block (fast_inflate path): the scanner thread checks their CRC-32 / length
struct MemberEnd { size_t end; uint32_t crc, isize; };
struct RawBlock { std::shared_ptr<TextBuf> buf; std::vector<MemberEnd> ends; bool crc_on = false; };
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

bool g_have_avx2 = false;
struct CpuInit { CpuInit() {
#if defined(__x86_64__)
    g_have_avx2 = __builtin_cpu_supports("avx2") && !getenv("MLGI_NO_AVX2");
#endif
} } g_cpu_init;
#if defined(__x86_64__)
// calls f(e) for the position e of every '\n' in base[0, n & ~31), in order; returns n & ~31
template <typename F>
__attribute__((target("avx2"))) size_t scan_newlines_avx2(const char* base, size_t n, F f) {
    const __m256i nl = _mm256_set1_epi8('\n');
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        uint32_t m = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_loadu_si256((const __m256i*)(base + i)), nl));
        while (m) { f(i + (size_t)__builtin_ctz(m)); m &= m - 1; }
    }
    return i;
}
__attribute__((target("avx2,popcnt"))) size_t count_newlines_avx2(const char* base, size_t n) {
    const __m256i nl = _mm256_set1_epi8('\n');
    size_t i = 0, c = 0;
    for (; i + 32 <= n; i += 32)
        c += (size_t)__builtin_popcount((uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_loadu_si256((const __m256i*)(base + i)), nl)));
    for (; i < n; ++i) c += base[i] == '\n';
    return c;
}
#endif
size_t count_newlines(const char* base, size_t n) {
#if defined(__x86_64__)
    if (g_have_avx2) return count_newlines_avx2(base, n);
#endif
    size_t c = 0;
    for (const char* q = base; (q = (const char*)memchr(q, '\n', (size_t)(base + n - q))) != nullptr; ++q) ++c;
    return c;
}

}  // namespace

struct mlgi_reader {
    gzFile fh = nullptr;
    int fd = -1;                        // plain (not gzip) files are read with read(2): no inflate layer in the way
    int type = MLGI_FASTQ;
    int threads = 1;
    size_t block_bytes = 8u << 20;
    std::shared_ptr<BufPool> pool = std::make_shared<BufPool>();
    BoundedQueue<RawBlock> q_raw{4};
    BoundedQueue<Lines> q_lines{4};
    std::thread t_read, t_scan;
    std::vector<uint32_t> spill;        // N runs of the last batch when they did not fit the caller's buffer
    std::string io_error;               // set by the reader thread before it finishes the queue
    std::mutex err_mu;
    // consumer state
    Lines cur;
    size_t cur_i = 0;
    bool eof = false;
    uint64_t tot_reads = 0, tot_bases = 0, tot_text = 0;
    double t_wait_lines = 0, t_gather = 0, t_pack = 0, t_scan_wait = 0;   // MLGI_PROFILE=1

    // plain gzip (one long deflate stream): fast_inflate.h on the mapped file, one thread -- the stream is sequential.
    // Every block starts HEAD bytes into its buffer; the last 32 KiB of the previous block are copied in front of it (the
    // decoder's match window, and at the same time the bytes the scanner would prepend as the carried-over line).
    void read_loop_gz_fast() {
        auto fail = [&](const char* m) { std::lock_guard<std::mutex> g(err_mu); io_error = m; q_raw.finish(); };
        fastinf::Inflater* z = new fastinf::Inflater();
        std::unique_ptr<fastinf::Inflater> zguard(z);
        z->in = (const uint8_t*)gzmap->p; z->in_end = z->in + gzmap->n;
        std::vector<uint8_t> hist;                       // tail of the previous block
        uint64_t member_bytes = 0;
        pool->bytes = HEAD + block_bytes + fastinf::SLACK + 8;
        for (bool more = true; more;) {
            RawBlock b;
            b.buf = std::make_shared<TextBuf>();
            b.buf->mem = pool->get();
            b.buf->pool = pool;
            uint8_t* start = (uint8_t*)b.buf->mem.get() + HEAD;
            if (!hist.empty()) memcpy(start - hist.size(), hist.data(), hist.size());
            const uint8_t* begin = start - hist.size();
            uint8_t* out = start;
            uint8_t* const limit = start + block_bytes;
            for (;;) {
                uint8_t* const seg = out;
                const int rc = z->run(out, limit, begin);
                if (rc < 0) { fail(z->error ? z->error : "corrupt gzip stream"); return; }
                member_bytes += (uint64_t)(out - seg);
                if (rc == 1) {                           // a member ended (the scanner thread checks its CRC); go on in the same block
                    if ((uint32_t)member_bytes != z->isize_expected) { fail("gzip length mismatch"); return; }
                    b.ends.push_back(MemberEnd{(size_t)(out - start), z->crc_expected, z->isize_expected});
                    member_bytes = 0;
                    if (out < limit) continue;
                    break;
                }
                if (rc == 2) { more = false; if (member_bytes) { fail("gzip stream ends inside a member"); return; } }
                break;                                   // rc == 0: the block is full
            }
            b.buf->n = (size_t)(out - start);
            b.crc_on = true;
            const size_t keep = std::min<size_t>((size_t)(out - begin), fastinf::WINDOW);
            hist.assign(out - keep, out);
            if (b.buf->n && !q_raw.push(std::move(b))) return;
        }
        q_raw.finish();
    }

    // BGZF (bgzip / htslib): a gzip file made of independent members of at most 64 KiB, each carrying its compressed size
    // in a 'BC' extra field and its inflated size in its trailer -- so members can be found without inflating anything and
    // inflated in parallel into known places.  Jobs of ~block_bytes of output are inflated `threads` at a time and handed
    // to the scanner in file order.
    std::shared_ptr<Mapping> gzmap;      // the compressed file, mapped (BGZF only)
    static bool bgzf_member(const unsigned char* p, size_t left, size_t* csize, size_t* hdr) {
        if (left < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return false;
        const size_t xlen = p[10] | ((size_t)p[11] << 8);
        if (left < 12 + xlen) return false;
        for (size_t o = 12; o + 4 <= 12 + xlen;) {
            const size_t slen = p[o + 2] | ((size_t)p[o + 3] << 8);
            if (p[o] == 'B' && p[o + 1] == 'C' && slen == 2 && o + 6 <= 12 + xlen) {
                *csize = (size_t)(p[o + 4] | ((size_t)p[o + 5] << 8)) + 1;
                *hdr = 12 + xlen;                                    // (FNAME / FCOMMENT / FHCRC are not used by BGZF writers)
                return *csize >= *hdr + 8 && *csize <= left;
            }
            o += 4 + slen;
        }
        return false;
    }
    void read_loop_bgzf() {
        const unsigned char* base = (const unsigned char*)gzmap->p;
        const size_t n = gzmap->n;
        struct Job { size_t in0, in1, out; std::shared_ptr<TextBuf> buf; bool ok = true; };
        const size_t P = (size_t)std::max(1, std::min(threads, 32));
        size_t pos = 0;
        auto fail = [&](const char* m) { std::lock_guard<std::mutex> g(err_mu); io_error = m; q_raw.finish(); };
        while (pos < n) {
            std::vector<Job> jobs;
            while (jobs.size() < P && pos < n) {                     // cut the next jobs: whole members, ~block_bytes of output each
                Job j; j.in0 = pos; j.out = 0;
                while (pos < n && j.out < block_bytes) {
                    size_t cs = 0, hd = 0;
                    if (!bgzf_member(base + pos, n - pos, &cs, &hd)) { fail("not a BGZF member where one was expected"); return; }
                    const unsigned char* t = base + pos + cs - 4;
                    j.out += (size_t)t[0] | ((size_t)t[1] << 8) | ((size_t)t[2] << 16) | ((size_t)t[3] << 24);
                    pos += cs;
                }
                j.in1 = pos;
                jobs.push_back(std::move(j));
            }
            auto inflate_job = [&](Job& j) {
                j.buf = std::make_shared<TextBuf>();
                j.buf->mem.reset(new char[HEAD + j.out + 1]);
                char* dst = j.buf->mem.get() + HEAD;
                size_t o = 0;
                for (size_t q = j.in0; q < j.in1;) {
                    size_t cs = 0, hd = 0;
                    bgzf_member(base + q, n - q, &cs, &hd);
                    const unsigned char* t = base + q + cs - 4;
                    const size_t isz = (size_t)t[0] | ((size_t)t[1] << 8) | ((size_t)t[2] << 16) | ((size_t)t[3] << 24);
                    z_stream zs;
                    memset(&zs, 0, sizeof(zs));
                    if (inflateInit2(&zs, -15) != Z_OK) { j.ok = false; return; }
                    zs.next_in = const_cast<unsigned char*>(base + q + hd); zs.avail_in = (uInt)(cs - hd - 8);
                    zs.next_out = (unsigned char*)dst + o; zs.avail_out = (uInt)isz;
                    const int rc = inflate(&zs, Z_FINISH);
                    bool good = rc == Z_STREAM_END && zs.total_out == isz;
                    inflateEnd(&zs);
                    if (good) {
                        const unsigned char* c = t - 4;
                        const unsigned long want = (unsigned long)c[0] | ((unsigned long)c[1] << 8) | ((unsigned long)c[2] << 16) | ((unsigned long)c[3] << 24);
                        good = crc32(crc32(0L, Z_NULL, 0), (const Bytef*)dst + o, (uInt)isz) == want;
                    }
                    if (!good) { j.ok = false; return; }
                    o += isz; q += cs;
                }
                j.buf->n = o;
            };
            std::vector<std::thread> th;
            for (size_t i = 1; i < jobs.size(); ++i) th.emplace_back([&, i] { inflate_job(jobs[i]); });
            inflate_job(jobs[0]);
            for (auto& x : th) x.join();
            for (auto& j : jobs) {
                if (!j.ok) { fail("corrupt BGZF member"); return; }
                RawBlock b; b.buf = j.buf;
                if (j.buf->n && !q_raw.push(std::move(b))) return;
            }
        }
        q_raw.finish();
    }

    void read_loop() {
        for (;;) {
            RawBlock b;
            b.buf = std::make_shared<TextBuf>();
            pool->bytes = HEAD + block_bytes;
            b.buf->mem = pool->get();
            b.buf->pool = pool;
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

    uint64_t line_no = 0;                // lines completed so far (FASTQ: record = 4 lines); scanner thread only
    std::shared_ptr<Mapping> map;        // plain files
    // cuts text into lines, appends the sequence lines to out; returns the bytes consumed (an incomplete last line is not)
    size_t emit_lines(const std::shared_ptr<TextBuf>& text, bool final_block, Lines& out) {
        const char* base = text->data();
        const size_t n = text->size();
        size_t p = 0;                                     // start of the current line
        out.start.reserve(n / 128); out.len.reserve(n / 128);
        auto line_ends_at = [&](size_t e) {               // the line [p, e) is complete (e = its '\n', or the end of the file)
            size_t le = e;
            if (le > p && base[le - 1] == '\r') --le;
            bool is_seq;
            if (type == MLGI_FASTQ) is_seq = (line_no & 3u) == 1u;
            else is_seq = le > p && base[p] != '>' && base[p] != ';';
            if (is_seq) { out.start.push_back((uint32_t)p); out.len.push_back((uint32_t)(le - p)); }
            ++line_no;
        };
        size_t q = 0;                                     // bytes [0, q) have been searched for '\n'
#if defined(__x86_64__)
        if (g_have_avx2) q = scan_newlines_avx2(base, n, [&](size_t e) { line_ends_at(e); p = e + 1; });
#endif
        while (q < n) {
            const char* nl = (const char*)memchr(base + q, '\n', n - q);
            if (!nl) break;
            const size_t e = (size_t)(nl - base);
            line_ends_at(e);
            p = e + 1; q = e + 1;
        }
        if (p < n && final_block) { line_ends_at(n); p = n; }   // the file's last line has no '\n'
        return p;                                         // bytes consumed; an incomplete line is carried into the next block
    }

    // One part of a parallel scan of the view vbase[0, vend): the lines that START in [b0, b1).  nl_before = lines
    // completed before b0 (= '\n' seen so far), so every part knows its lines' numbers without waiting for the others.
    // A line may end beyond b1; one that does not end before vend is complete only in the file's last view, otherwise
    // *open_line is set to its start.  Sequence lines go to out (offsets relative to vbase).
    void scan_part(const char* vbase, size_t b0, size_t b1, size_t vend, bool final_view, bool starts_line, uint64_t nl_before,
                   Lines& out, size_t* open_line) const {
        size_t p = b0;
        uint64_t idx = nl_before;
        if (!starts_line) {                                   // b0 is inside a line that belongs to an earlier part
            const char* nl = (const char*)memchr(vbase + b0, '\n', b1 - b0);
            if (!nl) return;                                  // no line starts in this part
            p = (size_t)(nl - vbase) + 1; ++idx;
        }
        out.start.reserve((b1 - b0) / 128 + 4); out.len.reserve((b1 - b0) / 128 + 4);
        auto line = [&](size_t e) {                           // the line [p, e) is complete
            size_t le = e;
            if (le > p && vbase[le - 1] == '\r') --le;
            bool is_seq;
            if (type == MLGI_FASTQ) is_seq = (idx & 3u) == 1u;
            else is_seq = le > p && vbase[p] != '>' && vbase[p] != ';';
            if (is_seq) { out.start.push_back((uint32_t)p); out.len.push_back((uint32_t)(le - p)); }
            ++idx;
            p = e + 1;
        };
        size_t q = p;                                         // bytes [p, q) hold no '\n'
#if defined(__x86_64__)
        if (g_have_avx2 && b1 > p) {
            const size_t from = p;
            q = from + scan_newlines_avx2(vbase + from, b1 - from, [&](size_t e) { line(from + e); });
        }
#endif
        while (p < b1) {                                      // (also finishes the last line that starts before b1)
            if (q < p) q = p;
            const char* nl = (const char*)memchr(vbase + q, '\n', vend - q);
            if (!nl) {
                if (final_view && p < vend) line(vend);
                else if (p < vend) *open_line = p;
                return;
            }
            q = (size_t)(nl - vbase) + 1;
            line(q - 1);
        }
    }

    // plain file: walk the mapping in views of `threads` blocks, scanned by that many threads (pass 1 counts the '\n' of
    // every part, pass 2 cuts the lines); a view's unterminated last line simply starts the next view
    void scan_loop_mapped() {
        const char* base = (const char*)map->p;
        const size_t n = map->n;
        const size_t P = (size_t)std::max(1, std::min(threads, 16));
        size_t pos = 0, span = block_bytes * P;
        while (pos < n) {
            const size_t end = std::min(n, pos + span);
            const size_t vn = end - pos;
            if (vn >= 0xFFFFFFF0ull) { std::lock_guard<std::mutex> g(err_mu); io_error = "a single line exceeds 4 GiB"; break; }
            auto text = std::make_shared<TextBuf>();
            text->map = map; text->ext = base + pos; text->n = vn;
            const char* vb = base + pos;
            const size_t parts = (P > 1 && vn >= 64 * P) ? P : 1;
            std::vector<size_t> bnd(parts + 1), cnt(parts, 0), open(parts, (size_t)-1);
            for (size_t i = 0; i <= parts; ++i) bnd[i] = vn * i / parts;
            std::vector<Lines> outs(parts);
            auto run = [&](auto&& fn) {
                std::vector<std::thread> th;
                for (size_t i = 1; i < parts; ++i) th.emplace_back([&fn, i] { fn(i); });
                fn(0);
                for (auto& x : th) x.join();
            };
            if (parts > 1) run([&](size_t i) { cnt[i] = count_newlines(vb + bnd[i], bnd[i + 1] - bnd[i]); });
            std::vector<uint64_t> before(parts + 1, line_no);
            for (size_t i = 0; i < parts; ++i) before[i + 1] = before[i] + cnt[i];
            run([&](size_t i) {
                outs[i].text = text;
                scan_part(vb, bnd[i], bnd[i + 1], vn, end == n, i == 0 || vb[bnd[i] - 1] == '\n', before[i], outs[i], &open[i]);
            });
            size_t used = vn;                                 // everything, unless some line is still open at the end of the view
            for (size_t i = 0; i < parts; ++i) if (open[i] != (size_t)-1) { used = open[i]; break; }
            if (used == 0 && end < n) { span += block_bytes * P; continue; }      // one line longer than the view: widen it
            // lines completed in this view: its '\n' up to `used`, plus the file's unterminated last line
            line_no += (parts > 1 ? before[parts] - line_no : (uint64_t)count_newlines(vb, vn)) + ((end == n && vn && vb[vn - 1] != '\n') ? 1u : 0u);
            pos += used; span = block_bytes * P;
            for (size_t i = 0; i < parts; ++i)
                if (!outs[i].start.empty() && !q_lines.push(std::move(outs[i]))) return;
        }
        q_lines.finish();
    }

    void scan_loop() {
        unsigned long run_crc = crc32(0L, Z_NULL, 0);
        std::string carry;               // the unterminated tail of the previous block
        RawBlock b;
        bool more = true;
        while (more) {
            const auto ts0 = std::chrono::steady_clock::now();
            more = q_raw.pop(b);
            t_scan_wait += std::chrono::duration<double>(std::chrono::steady_clock::now() - ts0).count();
            if (more && b.crc_on) {                           // gzip members decoded by fast_inflate: CRC-32 of their bytes
                const unsigned char* d = (const unsigned char*)b.buf->data();
                size_t p = 0;
                bool bad = false;
                for (const MemberEnd& m : b.ends) {
                    run_crc = crc32(run_crc, d + p, (uInt)(m.end - p));
                    if ((uint32_t)run_crc != m.crc) bad = true;
                    run_crc = crc32(0L, Z_NULL, 0);
                    p = m.end;
                }
                run_crc = crc32(run_crc, d + p, (uInt)(b.buf->size() - p));
                if (bad) { std::lock_guard<std::mutex> g(err_mu); io_error = "gzip CRC mismatch"; break; }
            }
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

// ---- packing: ASCII -> 2 bits per base, first base in the top bits of each byte ---------------------------------------
// code of a base from its ASCII byte, either case: (c >> 1) & 3 gives A=0 C=1 G=3 T=2; x ^ (x >> 1) swaps the last two
inline uint32_t codes4(uint32_t w) {
    uint32_t x = (w >> 1) & 0x03030303u;
    return x ^ ((x >> 1) & 0x01010101u);
}
// four codes (one per byte, first base in the low byte) -> one packed byte: the multiplier lines the 2-bit codes up in
// bits 24..31 of the product
inline uint8_t pack4(uint32_t codes) { return (uint8_t)((codes * 0x40100401u) >> 24); }
// 0x80 in every byte of x that is NOT one of a/c/g/t (x already has bit 5 set in every byte)
inline uint64_t not_acgt8(uint64_t x) {
    const uint64_t L7 = 0x7F7F7F7F7F7F7F7Full;
    auto nz = [&](uint64_t z) { return (((z & L7) + L7) | z); };          // bit 7 of every non-zero byte
    return nz(x ^ 0x6161616161616161ull) & nz(x ^ 0x6363636363636363ull) & nz(x ^ 0x6767676767676767ull) &
           nz(x ^ 0x7474747474747474ull) & 0x8080808080808080ull;
}
// n bases (n % 4 == 0), all of them A/C/G/T in either case, to n / 4 output bytes; returns false (nothing written that
// matters) at the first 8-byte group that holds another symbol
#if defined(__x86_64__)
__attribute__((target("avx2"))) size_t pack_clean_avx2(const unsigned char* s, size_t n, uint8_t* out) {
    const __m256i lo5 = _mm256_set1_epi8((char)0xDF);
    const __m256i cA = _mm256_set1_epi8('A'), cC = _mm256_set1_epi8('C'), cG = _mm256_set1_epi8('G'), cT = _mm256_set1_epi8('T');
    const __m256i three = _mm256_set1_epi8(3), one = _mm256_set1_epi8(1);
    const __m256i m41 = _mm256_set1_epi16(0x0104);          // bytes (4, 1): first base of a pair times 4 plus the second
    const __m256i m161 = _mm256_set1_epi32(0x00010010);     // words (16, 1)
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        const __m256i v = _mm256_loadu_si256((const __m256i*)(s + i));
        const __m256i u = _mm256_and_si256(v, lo5);
        const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(u, cA), _mm256_cmpeq_epi8(u, cC)),
                                           _mm256_or_si256(_mm256_cmpeq_epi8(u, cG), _mm256_cmpeq_epi8(u, cT)));
        if (_mm256_movemask_epi8(ok) != -1) break;
        __m256i x = _mm256_and_si256(_mm256_srli_epi16(v, 1), three);
        x = _mm256_xor_si256(x, _mm256_and_si256(_mm256_srli_epi16(x, 1), one));
        const __m256i p2 = _mm256_maddubs_epi16(x, m41);      // 16-bit: c0 * 4 + c1
        const __m256i p4 = _mm256_madd_epi16(p2, m161);       // 32-bit: (c0 * 4 + c1) * 16 + (c2 * 4 + c3)
        const __m256i w16 = _mm256_packus_epi32(p4, p4);      // per 128-bit lane: 4 values as 16-bit, twice
        const __m256i w8 = _mm256_packus_epi16(w16, w16);     // per lane: 4 bytes, four times
        const uint32_t a = (uint32_t)_mm256_extract_epi32(w8, 0), b = (uint32_t)_mm256_extract_epi32(w8, 4);
        memcpy(out + i / 4, &a, 4);
        memcpy(out + i / 4 + 4, &b, 4);
    }
    return i;
}
#endif
size_t pack_clean_swar(const unsigned char* s, size_t n, uint8_t* out) {
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t v;
        memcpy(&v, s + i, 8);
        if (not_acgt8(v | 0x2020202020202020ull)) break;
        out[i / 4] = pack4(codes4((uint32_t)v));
        out[i / 4 + 1] = pack4(codes4((uint32_t)(v >> 32)));
    }
    return i;
}
inline size_t pack_clean(const unsigned char* s, size_t n, uint8_t* out) {
#if defined(__x86_64__)
    if (g_have_avx2) { const size_t d = pack_clean_avx2(s, n, out); return d + pack_clean_swar(s + d, n - d, out + d / 4); }
#endif
    return pack_clean_swar(s, n, out);
}

void pack_range(PackJob& j) {
    if (j.a >= j.b) return;
    uint64_t pos = j.off[j.a];
    uint8_t* out = j.bases + (pos >> 2);
    unsigned fill = (unsigned)(pos & 3u);      // bases already present in the current output byte (owned by a neighbour)
    unsigned acc = 0;
    bool first_partial = fill != 0;
    uint64_t run_start = 0; uint32_t run_len = 0;
    auto one = [&](unsigned char ch) {         // the general path: one base, N bookkeeping, byte assembly
        unsigned c = g_code[ch];
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
    };
    for (size_t i = j.a; i < j.b; ++i) {
        const unsigned char* s = (const unsigned char*)j.ptr[i];
        const uint32_t L = j.len[i];
        uint32_t k = 0;
        while (k < L) {
            // whole output bytes of clean bases go through the vector / SWAR path; everything else one base at a time
            if (fill == 0 && L - k >= 8) {
                const size_t d = pack_clean(s + k, (size_t)((L - k) & ~3u), out);
                out += d / 4; pos += d; k += (uint32_t)d;
                if (k >= L) break;
                // an unclean group (or the read's tail) follows: take up to 8 bases the slow way, then try again
                const uint32_t lim = std::min<uint32_t>(L, k + 8);
                while (k < lim) one(s[k++]);
            } else {
                one(s[k++]);
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
    std::shared_ptr<Mapping> gzmap;
    if (gz && !getenv("MLGI_NO_BGZF")) {           // BGZF? then the members are inflated in parallel from a mapping of the file
        unsigned char h[64];
        const ssize_t got = pread(fd, h, sizeof(h), 0);
        size_t cs = 0, hd = 0;
        struct stat sb;
        if (got >= 18 && fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size >= 28) {
            // (the member may be longer than the 64 bytes peeked at: only its header has to be inside them)
            const size_t xlen = h[10] | ((size_t)h[11] << 8);
            if (12 + xlen <= (size_t)got && mlgi_reader::bgzf_member(h, (size_t)sb.st_size, &cs, &hd)) {
                void* p = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
                if (p != MAP_FAILED) { gzmap = std::make_shared<Mapping>(); gzmap->p = p; gzmap->n = (size_t)sb.st_size; }
            }
        }
    }
    bool fast_gz = false;
    if (gz && !gzmap && !getenv("MLGI_ZLIB")) {        // plain gzip: map it for the in-house decoder
        struct stat sb;
        if (fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size >= 20) {
            void* p = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (p != MAP_FAILED) { gzmap = std::make_shared<Mapping>(); gzmap->p = p; gzmap->n = (size_t)sb.st_size; fast_gz = true; }
        }
    }
    if (gz && !gzmap) {
        fh = gzdopen(fd, "rb");
        if (!fh) { close(fd); set_error("cannot open %s", path); return -3; }
        gzbuffer(fh, 1u << 20);
        fd = -1;
    }
#if defined(__x86_64__)
    g_have_avx2 = __builtin_cpu_supports("avx2") && !getenv("MLGI_NO_AVX2");     // (the switch is for tests)
#endif
    mlgi_reader* r = new mlgi_reader();
    r->fh = fh; r->fd = fd; r->type = input_type; r->gzmap = gzmap;
    int hw = (int)std::thread::hardware_concurrency();
    if (hw < 1) hw = 1;
    r->threads = threads > 0 ? threads : hw;
    if (const char* s = getenv("MLGI_BLOCK_BYTES")) { long v = atol(s); if (v >= 16) r->block_bytes = (size_t)v; }
    if (fd >= 0 && !gz && !getenv("MLGI_NO_MMAP")) {
        struct stat sb;
        if (fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size > 0) {
            void* p = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (p != MAP_FAILED) {
                // (no MADV_SEQUENTIAL: it lets the kernel drop the pages right behind the scan, and a file that is read
                // again -- paired runs, several k ranges -- would come from the disk instead of the page cache)
                r->map = std::make_shared<Mapping>();
                r->map->p = p; r->map->n = (size_t)sb.st_size;
            }
        }
    }
    if (r->map) r->t_scan = std::thread([r] { r->scan_loop_mapped(); });
    else {
        if (r->gzmap && fast_gz) r->t_read = std::thread([r] { r->read_loop_gz_fast(); });
        else if (r->gzmap) r->t_read = std::thread([r] { r->read_loop_bgzf(); });
        else r->t_read = std::thread([r] { r->read_loop(); });
        r->t_scan = std::thread([r] { r->scan_loop(); });
    }
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
            const auto tw0 = std::chrono::steady_clock::now();
            const bool got_lines = r->q_lines.pop(nx);
            r->t_wait_lines += std::chrono::duration<double>(std::chrono::steady_clock::now() - tw0).count();
            if (!got_lines) { r->eof = true; break; }
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
    const auto tg0 = std::chrono::steady_clock::now();
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
    const auto tp0 = std::chrono::steady_clock::now();
    r->t_gather += std::chrono::duration<double>(tp0 - tg0).count();
    if (T == 1) pack_range(jobs[0]);
    else {
        std::vector<std::thread> th;
        for (int t = 1; t < T; ++t) th.emplace_back([&jobs, t] { pack_range(jobs[(size_t)t]); });
        pack_range(jobs[0]);
        for (auto& x : th) x.join();
    }
    r->t_pack += std::chrono::duration<double>(std::chrono::steady_clock::now() - tp0).count();
    // N runs of the workers, in stream order; runs that touch across a worker boundary are merged
    // A batch with more runs than the caller's buffer holds (very low-quality reads: KMC in the reference just skips N,
    // select_db.py:50) is still delivered: its runs stay with the reader and are fetched with mlgi_spilled_runs().
    uint64_t nr = 0, upper = 0;
    for (int t = 0; t < T; ++t) upper += jobs[(size_t)t].runs.size() / 2;
    const bool spill = upper > cap_runs || !nruns;
    if (spill) r->spill.assign((size_t)(2 * upper + 2), 0u);
    uint32_t* dst = spill ? r->spill.data() : nruns;
    for (int t = 0; t < T; ++t) {
        const std::vector<uint32_t>& v = jobs[(size_t)t].runs;
        for (size_t i = 0; i + 1 < v.size(); i += 2) {
            if (nr && (uint64_t)dst[2 * (nr - 1)] + dst[2 * (nr - 1) + 1] == v[i]) { dst[2 * (nr - 1) + 1] += v[i + 1]; continue; }
            dst[2 * nr] = v[i]; dst[2 * nr + 1] = v[i + 1];
            ++nr;
        }
    }
    if (spill) r->spill.resize((size_t)(2 * nr));
    *n_reads = n; *n_runs = nr;
    r->tot_reads += n; r->tot_bases += nb;
    return spill && nr ? 2 : 1;
}

MLGI_API int mlgi_spilled_runs(mlgi_reader* r, uint32_t* nruns, uint64_t cap_runs) {
    if (!r || !nruns) { set_error("null argument"); return -2; }
    if (cap_runs * 2 < r->spill.size()) { set_error("buffer holds %llu runs, %llu are waiting", (unsigned long long)cap_runs, (unsigned long long)(r->spill.size() / 2)); return -2; }
    if (!r->spill.empty()) memcpy(nruns, r->spill.data(), r->spill.size() * sizeof(uint32_t));
    r->spill.clear();
    return 0;
}

MLGI_API int mlgi_stats(mlgi_reader* r, uint64_t* reads, uint64_t* bases, uint64_t* text_bytes) {
    if (!r) { set_error("null reader"); return -2; }
    if (reads) *reads = r->tot_reads;
    if (bases) *bases = r->tot_bases;
    if (text_bytes) *text_bytes = r->tot_text;
    return 0;
}

// test hook (not in the header): gzip-decode in[0, n) with fast_inflate.h into out[0, cap), stopping and resuming every
// `step` output bytes.  Returns 0 and *out_n on success, -1 on a decoding error, -2 if cap is too small, -3 on a CRC /
// length mismatch.
MLGI_API int mlgi_test_inflate(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap, uint64_t step, uint64_t* out_n) {
    std::unique_ptr<fastinf::Inflater> z(new fastinf::Inflater());
    z->in = in; z->in_end = in + n;
    uint8_t* o = out;
    uint8_t* member = out;
    for (;;) {
        if ((uint64_t)(o - out) + step + fastinf::SLACK > cap) return -2;
        const int rc = z->run(o, o + step, out);
        if (rc < 0) { set_error("%s", z->error ? z->error : "?"); return -1; }
        if (rc == 1) {
            if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), member, (uInt)(o - member)) != z->crc_expected || (uint32_t)(o - member) != z->isize_expected) return -3;
            member = o;
        }
        if (rc == 2) break;
    }
    *out_n = (uint64_t)(o - out);
    return 0;
}

// ---- KMC database reader (kmcdb.h) ----
struct mlgi_kmcdb { kmcdb::Reader r; };
MLGI_API int mlgi_kmc_open(const char* prefix, mlgi_kmcdb** out) {
    if (!prefix || !out) { set_error("null argument"); return -2; }
    std::unique_ptr<mlgi_kmcdb> d(new mlgi_kmcdb());
    if (!d->r.open(prefix)) { set_error("%s", d->r.error.c_str()); return -3; }
    *out = d.release();
    return 0;
}
MLGI_API int mlgi_kmc_info(mlgi_kmcdb* d, uint32_t* k, uint64_t* total, uint32_t* counter_size, uint32_t* min_count, uint64_t* max_count,
                           int* canonical, uint32_t* version) {
    if (!d) { set_error("null database"); return -2; }
    const kmcdb::Info& i = d->r.info;
    if (k) *k = i.k;
    if (total) *total = i.total;
    if (counter_size) *counter_size = i.counter_size;
    if (min_count) *min_count = i.min_count;
    if (max_count) *max_count = i.max_count;
    if (canonical) *canonical = i.canonical ? 1 : 0;
    if (version) *version = i.version;
    return 0;
}
MLGI_API int mlgi_kmc_read(mlgi_kmcdb* d, uint64_t* keys, uint32_t* counts_or_null) {
    if (!d || !keys) { set_error("null argument"); return -2; }
    if (!d->r.read_all(keys, counts_or_null)) { set_error("%s", d->r.error.c_str()); return -3; }
    return 0;
}
MLGI_API void mlgi_kmc_close(mlgi_kmcdb* d) { delete d; }

MLGI_API void mlgi_close(mlgi_reader* r) {
    if (!r) return;
    if (getenv("MLGI_PROFILE"))
        fprintf(stderr, "mlgi: consumer waited %.3f s for lines, prepared %.3f s, packed %.3f s; scanner waited %.3f s for blocks\n",
                r->t_wait_lines, r->t_gather, r->t_pack, r->t_scan_wait);
    r->q_lines.abandon();
    r->q_raw.abandon();
    if (r->t_scan.joinable()) r->t_scan.join();
    if (r->t_read.joinable()) r->t_read.join();
    if (r->fh) gzclose(r->fh);
    if (r->fd >= 0) close(r->fd);
    delete r;
}

}  // extern "C"
