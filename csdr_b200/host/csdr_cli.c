/*
 * csdr_cli.c -- the `csdr` stdin/stdout command surface for the hot-path commands, host code in C over
 * libcsdr_b200.so (Part A of include/csdr_b200.h).  One process = one DSP block, raw samples on the pipes,
 * exactly like the reference CLI, so existing pipe graphs (csdr-fm:41, README.md:54-110) drop in unchanged.
 *
 * Behaviours reproduced from the reference (all citations into reference csdr.c):
 *   - block sizes: 1024 default / 16384 for the wideband commands, rounded up to a multiple of 4 (:189-190,
 *     :353-357), fir_decimate_cc grows its block to >= 2*taps (:1136); CSDR_FIXED_BUFSIZE,
 *     CSDR_DYNAMIC_BUFSIZE_ON, CSDR_PRINT_BUFSIZES (:394-417)
 *   - optional 8-byte "csdr"+int preamble in and out when dynamic buffer sizes are on (:330-339, :377-392)
 *   - EOF framing: the end-of-file test comes BEFORE the read, so a short final read is still processed and
 *     written as a whole block (:248 and every loop); shift_addition_cc alone stops on an empty read (:907)
 *   - fflush + sched_yield after every block (:198); F_SETPIPE_SZ 2 MiB, 4096 for small blocks (:427-428, :369-373)
 *   - runtime retune over --fifo <path> / --fd <n>: text lines, non-blocking, last complete line wins (:252-323)
 *   - the stderr lines the reference prints for these commands
 * Everything else the reference CLI offers (100+ other commands) is out of scope.
 */
#define _GNU_SOURCE
#include "csdr_b200.h"

#include <fcntl.h>
#include <math.h>
#include <sched.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <unistd.h>

/* ---- process-wide settings ---------------------------------------------------------------------- */
static struct {
    int fixed_small, fixed_big, dynamic_on, print_sizes, wideband;
    int argc; char **argv;
} G = {1024, 1024 * 16, 0, 0, 0, 0, NULL};

static void who(void) { fprintf(stderr, "%s%s%s: ", G.argv[0], G.argc >= 2 ? " " : "", G.argc >= 2 ? G.argv[1] : ""); }

static int complain(const char *why) { who(); fprintf(stderr, "%s\n", why); return -1; }

static void read_environment(void)
{
    const char *v = getenv("CSDR_DYNAMIC_BUFSIZE_ON");
    if (v) { G.dynamic_on = !!atoi(v); G.fixed_small = 0; }
    else if ((v = getenv("CSDR_FIXED_BUFSIZE"))) G.fixed_big = G.fixed_small = atoi(v);
    if ((v = getenv("CSDR_PRINT_BUFSIZES"))) G.print_sizes = atoi(v);
}

/* ---- buffer-size negotiation ------------------------------------------------------------------- */
static const char kPreamble[4] = {'c', 's', 'd', 'r'};

static int incoming_block_size(void)
{
    if (!G.dynamic_on) return G.wideband ? G.fixed_big : G.fixed_small;
    int head[2] = {0, 0};
    if (fread(head, sizeof(int), 2, stdin) != 2 || memcmp(head, kPreamble, 4) != 0) {
        complain("warning! Did not match preamble on the beginning of the stream. You should put \"csdr setbuf <buffer size>\" at the "
                 "beginning of the chain! Falling back to default buffer size: 1024");
        return 1024;
    }
    if (head[1] <= 0) { complain("warning! Invalid buffer size."); return 0; }
    return head[1];
}

static int round_to_unit(int n) { return n <= 0 ? 4 : ((n - 1) & ~3) + 4; }

static int block = 0;                                  /* samples (or values) per block of this process */

static int open_block(void)
{
    block = incoming_block_size();
    if (!block) return 0;
    block = round_to_unit(block);
    if (G.print_sizes) { who(); fprintf(stderr, "buffer size set to %d\n", block); }
    if (block <= 4096) { fcntl(STDIN_FILENO, F_SETPIPE_SZ, 4096); fcntl(STDOUT_FILENO, F_SETPIPE_SZ, 4096); }
    return block;
}

static int announce_block(int size)
{
    if (size <= 4096) fcntl(STDOUT_FILENO, F_SETPIPE_SZ, 4096);
    if (!G.dynamic_on) return G.fixed_small;
    if (G.print_sizes) { who(); fprintf(stderr, "next process proposed input buffer size is %d\n", size); }
    int head[2]; memcpy(head, kPreamble, 4); head[1] = size;
    fwrite(head, sizeof(int), 2, stdout);
    return size;
}

static void end_of_block(void) { fflush(stdout); sched_yield(); }

static void *must_alloc(size_t bytes) { void *p = calloc(1, bytes ? bytes : 1); if (!p) { complain("out of memory"); exit(-2); } return p; }

/* ---- runtime control channel (--fifo / --fd) ----------------------------------------------------- */
static int open_control(int argc, char **argv)
{
    if (argc < 4) return 0;
    int fd = 0;
    if (!strcmp(argv[2], "--fifo")) { who(); fprintf(stderr, "fifo control mode on\n"); fd = open(argv[3], O_RDONLY); }
    else if (!strcmp(argv[2], "--fd")) { if (sscanf(argv[3], "%d", &fd) <= 0) return 0; who(); fprintf(stderr, "fd control mode on, fd=%d\n", fd); }
    else return 0;
    fcntl(fd, F_SETFL, fcntl(fd, F_GETFL, 0) | O_NONBLOCK);
    return fd;
}

/* returns 1 when at least one complete line arrived; the LAST complete line is parsed with `format` */
static int poll_control(int fd, const char *format, ...)
{
    static char pending[1024];
    static int have = 0;
    if (!fd) return 0;
    int got = (int)read(fd, pending + have, sizeof pending - (size_t)have);
    if (got <= 0) return 0;
    int total = have + got, last_end = 0, prev_end = 0;
    for (int k = 0; k < total; k++) if (pending[k] == '\n') { prev_end = last_end; last_end = k + 1; }
    if (!last_end) { have = total; return 0; }
    va_list ap; va_start(ap, format); vsscanf(pending + prev_end, format, ap); va_end(ap);
    memmove(pending, pending + last_end, (size_t)(total - last_end));
    have = total - last_end;
    return 1;
}

/* ---- commands ------------------------------------------------------------------------------------ */
static int cmd_convert_u8_f(int argc, char **argv)
{
    (void)argc; (void)argv;
    if (!announce_block(open_block())) return -2;
    unsigned char *in = must_alloc((size_t)block);
    float *out = must_alloc(sizeof(float) * (size_t)block);
    for (;;) {
        if (feof(stdin)) return 0;
        fread(in, 1, (size_t)block, stdin);
        convert_u8_f(in, out, block);
        fwrite(out, sizeof(float), (size_t)block, stdout);
        end_of_block();
    }
}

static int cmd_convert_s16_f(int argc, char **argv)
{
    (void)argc; (void)argv;
    if (!announce_block(open_block())) return -2;
    short *in = must_alloc(sizeof(short) * (size_t)block);
    float *out = must_alloc(sizeof(float) * (size_t)block);
    for (;;) {
        if (feof(stdin)) return 0;
        fread(in, sizeof(short), (size_t)block, stdin);
        convert_s16_f(in, out, block);
        fwrite(out, sizeof(float), (size_t)block, stdout);
        end_of_block();
    }
}

static int cmd_convert_f_s16(int argc, char **argv)
{
    (void)argc; (void)argv;
    if (!announce_block(open_block())) return -2;
    float *in = must_alloc(sizeof(float) * (size_t)block);
    short *out = must_alloc(sizeof(short) * (size_t)block);
    for (;;) {
        if (feof(stdin)) return 0;
        fread(in, sizeof(float), (size_t)block, stdin);
        convert_f_s16(in, out, block);
        fwrite(out, sizeof(short), (size_t)block, stdout);
        end_of_block();
    }
}

static int cmd_shift_addition_cc(int argc, char **argv)
{
    G.wideband = 1;
    float phase = 0, rate = 0;
    int ctl = open_control(argc, argv);
    if (ctl) { while (!poll_control(ctl, "%g\n", &rate)) usleep(10000); }
    else { if (argc <= 2) return complain("need required parameter (rate)"); sscanf(argv[2], "%g", &rate); }
    if (!announce_block(open_block())) return -2;
    complexf *in = must_alloc(sizeof(complexf) * (size_t)block), *out = must_alloc(sizeof(complexf) * (size_t)block);
    for (;;) {
        shift_addition_data_t nco = shift_addition_init(rate);
        who(); fprintf(stderr, "reinitialized to %g\n", rate);
        for (;;) {
            if (feof(stdin)) return 0;
            if (!fread(in, sizeof(complexf), (size_t)block, stdin)) break;
            for (int done = 0; done < block;) {                         /* the phasor is re-seeded every <= 1024 samples (:911-918) */
                int n = block - done > 1024 ? 1024 : block - done;
                phase = shift_addition_cc(in + done, out + done, n, nco, phase);
                done += n;
            }
            fwrite(out, sizeof(complexf), (size_t)block, stdout);
            if (poll_control(ctl, "%g\n", &rate)) break;
            end_of_block();
        }
    }
}

static int cmd_fir_decimate_cc(int argc, char **argv)
{
    G.wideband = 1;
    if (argc <= 2) return complain("need required parameter (decimation factor)");
    int factor = 0; sscanf(argv[2], "%d", &factor);
    float transition_bw = 0.05f; if (argc >= 4) sscanf(argv[3], "%g", &transition_bw);
    window_t window = WINDOW_DEFAULT;
    if (argc >= 5) window = firdes_get_window_from_string(argv[4]);
    else fprintf(stderr, "fir_decimate_cc: window = %s\n", firdes_get_string_from_window(window));
    int taps_length = firdes_filter_len(transition_bw);
    fprintf(stderr, "fir_decimate_cc: taps_length = %d\n", taps_length);
    while (G.fixed_big < taps_length * 2) G.fixed_big *= 2;
    if (!open_block()) return -2;
    announce_block(block / factor);
    float *taps = must_alloc(sizeof(float) * (size_t)taps_length);
    firdes_lowpass_f(taps, taps_length, 0.5f / (float)factor, window);
    complexf *in = must_alloc(sizeof(complexf) * (size_t)block), *out = must_alloc(sizeof(complexf) * (size_t)block);
    fread(in, sizeof(complexf), (size_t)block, stdin);
    for (;;) {
        if (feof(stdin)) return 0;
        int produced = fir_decimate_cc(in, out, block, factor, taps, taps_length);
        fwrite(out, sizeof(complexf), (size_t)produced, stdout);
        end_of_block();
        int consumed = factor * produced;                                /* keep the unconsumed tail, refill behind it (:1172-1174) */
        memmove(in, in + consumed, sizeof(complexf) * (size_t)(block - consumed));
        fread(in + (block - consumed), sizeof(complexf), (size_t)consumed, stdin);
    }
}

static int cmd_fmdemod_quadri_cf(int argc, char **argv)
{
    (void)argc; (void)argv;
    if (!announce_block(open_block())) return -2;
    complexf *in = must_alloc(sizeof(complexf) * (size_t)block);
    float *out = must_alloc(sizeof(float) * (size_t)block);
    complexf last = {0.f, 0.f};
    for (;;) {
        if (feof(stdin)) return 0;
        fread(in, sizeof(complexf), (size_t)block, stdin);
        last = fmdemod_quadri_cf(in, out, block, NULL, last);
        fwrite(out, sizeof(float), (size_t)block, stdout);
        end_of_block();
    }
}

static int cmd_shift_unroll_cc(int argc, char **argv)                      /* csdr.c:800-849 */
{
    G.wideband = 1;
    float phase = 0, rate = 0;
    int ctl = open_control(argc, argv);
    if (ctl) { while (!poll_control(ctl, "%g\n", &rate)) usleep(10000); }
    else { if (argc <= 2) return complain("need required parameter (rate)"); sscanf(argv[2], "%g", &rate); }
    if (!announce_block(open_block())) return -2;
    complexf *in = must_alloc(sizeof(complexf) * (size_t)block), *out = must_alloc(sizeof(complexf) * (size_t)block);
    for (;;) {
        shift_unroll_data_t table = shift_unroll_init(rate, 1024);
        who(); fprintf(stderr, "reinitialized to %g\n", rate);
        for (;;) {
            if (feof(stdin)) return 0;
            if (!fread(in, sizeof(complexf), (size_t)block, stdin)) break;
            for (int done = 0; done < block;) {
                int n = block - done > 1024 ? 1024 : block - done;
                phase = shift_unroll_cc(in + done, out + done, n, &table, phase);
                done += n;
            }
            fwrite(out, sizeof(complexf), (size_t)block, stdout);
            if (poll_control(ctl, "%g\n", &rate)) break;
            end_of_block();
        }
        free(table.dsin); free(table.dcos);
    }
}

static int cmd_shift_math_cc(int argc, char **argv)                        /* csdr.c:703-718 */
{
    if (argc <= 2) return complain("need required parameter (rate)");
    float phase = 0, rate = 0; sscanf(argv[2], "%g", &rate);
    if (!announce_block(open_block())) return -2;
    complexf *in = must_alloc(sizeof(complexf) * (size_t)block), *out = must_alloc(sizeof(complexf) * (size_t)block);
    for (;;) {
        if (feof(stdin)) return 0;
        if (!fread(in, sizeof(complexf), (size_t)block, stdin)) return 0;
        phase = shift_math_cc(in, out, block, rate, phase);
        fwrite(out, sizeof(complexf), (size_t)block, stdout);
        end_of_block();
    }
}

static int cmd_shift_table_cc(int argc, char **argv)                       /* csdr.c:725-747 */
{
    G.wideband = 1;
    if (argc <= 2) return complain("need required parameter (rate)");
    float phase = 0, rate = 0; sscanf(argv[2], "%g", &rate);
    int table_size = 65536; if (argc > 3) sscanf(argv[3], "%d", &table_size);
    if (!announce_block(open_block())) return -2;
    shift_table_data_t table = shift_table_init(table_size);
    who(); fprintf(stderr, "LUT initialized\n");
    complexf *in = must_alloc(sizeof(complexf) * (size_t)block), *out = must_alloc(sizeof(complexf) * (size_t)block);
    for (;;) {
        if (feof(stdin)) return 0;
        if (!fread(in, sizeof(complexf), (size_t)block, stdin)) return 0;
        phase = shift_table_cc(in, out, block, rate, table, phase);
        fwrite(out, sizeof(complexf), (size_t)block, stdout);
        end_of_block();
    }
}

static int cmd_shift_addfast_cc(int argc, char **argv)                     /* csdr.c:749-798 */
{
    G.wideband = 1;
    float phase = 0, rate = 0;
    int ctl = open_control(argc, argv);
    if (ctl) { while (!poll_control(ctl, "%g\n", &rate)) usleep(10000); }
    else { if (argc <= 2) return complain("need required parameter (rate)"); sscanf(argv[2], "%g", &rate); }
    if (!announce_block(open_block())) return -2;
    complexf *in = must_alloc(sizeof(complexf) * (size_t)block), *out = must_alloc(sizeof(complexf) * (size_t)block);
    for (;;) {
        shift_addfast_data_t steps = shift_addfast_init(rate);
        who(); fprintf(stderr, "reinitialized to %g\n", rate);
        for (;;) {
            if (feof(stdin)) return 0;
            if (!fread(in, sizeof(complexf), (size_t)block, stdin)) break;
            for (int done = 0; done < block;) {
                int n = block - done > 1024 ? 1024 : block - done;
                phase = shift_addfast_cc(in + done, out + done, n, &steps, phase);
                done += n;
            }
            fwrite(out, sizeof(complexf), (size_t)block, stdout);
            if (poll_control(ctl, "%g\n", &rate)) break;
            end_of_block();
        }
    }
}

static int cmd_decimating_shift_addition_cc(int argc, char **argv)         /* csdr.c:851-875 */
{
    G.wideband = 1;
    if (argc <= 2) return complain("need required parameter (rate)");
    float rate = 0; sscanf(argv[2], "%g", &rate);
    int decimation = 1; if (argc > 3) sscanf(argv[3], "%d", &decimation);
    if (!open_block()) return -2;
    announce_block(block / decimation);
    shift_addition_data_t nco = decimating_shift_addition_init(rate, decimation);
    decimating_shift_addition_status_t st = {0, 0.f, 0};
    complexf *in = must_alloc(sizeof(complexf) * (size_t)block), *out = must_alloc(sizeof(complexf) * (size_t)block);
    for (;;) {
        if (feof(stdin)) return 0;
        if (!fread(in, sizeof(complexf), (size_t)block, stdin)) return 0;
        st = decimating_shift_addition_cc(in, out, block, nco, decimation, st);
        fwrite(out, sizeof(complexf), (size_t)st.output_size, stdout);
        end_of_block();
    }
}

static int cmd_fft_cc(int argc, char **argv)                               /* csdr.c:1569-1643 (binary output; --octave is not part of this build) */
{
    if (argc <= 3) return complain("need required parameters (fft_size, out_of_every_n_samples)");
    int fft_size = 0; sscanf(argv[2], "%d", &fft_size);
    if (log2n(fft_size) == -1) return complain("fft_size should be power of 2");
    int every = 0; sscanf(argv[3], "%d", &every);
    window_t window = WINDOW_DEFAULT;
    if (argc >= 5) window = firdes_get_window_from_string(argv[4]);
    if (!open_block()) return -2;
    announce_block(fft_size);
    complexf *in = fft_malloc(sizeof(complexf) * (size_t)fft_size), *win = fft_malloc(sizeof(complexf) * (size_t)fft_size);
    complexf *out = fft_malloc(sizeof(complexf) * (size_t)fft_size), *skip = must_alloc(sizeof(complexf) * (size_t)block);
    FFT_PLAN_T *plan = make_fft_c2c(fft_size, win, out, 1, 0);
    if (!plan) return complain("FFT size error.");
    float *table = precalculate_window(fft_size, window);
    memset(in, 0, sizeof(complexf) * (size_t)fft_size);
    for (;;) {
        if (feof(stdin)) return 0;
        if (every > fft_size) {
            fread(in, sizeof(complexf), (size_t)fft_size, stdin);
            for (int remain = every - fft_size; remain > 0; remain -= block) fread(skip, sizeof(complexf), (size_t)(remain < block ? remain : block), stdin);
        } else {
            memmove(in, in + every, sizeof(complexf) * (size_t)(fft_size - every));
            fread(in + fft_size - every, sizeof(complexf), (size_t)every, stdin);
        }
        apply_precalculated_window_c(in, win, fft_size, table);
        fft_execute(plan);
        fwrite(out, sizeof(complexf), (size_t)fft_size, stdout);
        end_of_block();
    }
}

static int cmd_logpower_cf(int argc, char **argv)                          /* csdr.c:1645-1661 */
{
    float add_db = 0; if (argc >= 3) sscanf(argv[2], "%g", &add_db);
    if (!announce_block(open_block())) return -2;
    complexf *in = must_alloc(sizeof(complexf) * (size_t)block);
    float *out = must_alloc(sizeof(float) * (size_t)block);
    for (;;) {
        if (feof(stdin)) return 0;
        fread(in, sizeof(complexf), (size_t)block, stdin);
        logpower_cf(in, out, block, add_db);
        fwrite(out, sizeof(float), (size_t)block, stdout);
        end_of_block();
    }
}

static int cmd_logaveragepower_cf(int argc, char **argv)                   /* csdr.c:1663-1695 */
{
    G.wideband = 1;
    if (argc <= 4) return complain("need required parameters (add_db, fft_size, avgnumber)");
    float add_db = 0; int fft_size = 0, avgnumber = 0;
    sscanf(argv[2], "%g", &add_db); sscanf(argv[3], "%d", &fft_size); sscanf(argv[4], "%d", &avgnumber);
    complexf *in = must_alloc(sizeof(complexf) * (size_t)fft_size);
    float *acc = must_alloc(sizeof(float) * (size_t)fft_size);
    add_db -= 10.0 * log10(avgnumber);
    for (;;) {
        memset(acc, 0, sizeof(float) * (size_t)fft_size);
        if (feof(stdin)) return 0;
        for (int n = 0; n < avgnumber; n++) { fread(in, sizeof(complexf), (size_t)fft_size, stdin); accumulate_power_cf(in, acc, fft_size); }
        log_ff(acc, acc, fft_size, add_db);
        fwrite(acc, sizeof(float), (size_t)fft_size, stdout);
        end_of_block();
    }
}

static int cmd_fft_exchange_sides_ff(int argc, char **argv)                /* csdr.c:1697-1715: pure I/O, the two halves of every line swap places */
{
    if (argc <= 2) return complain("need required parameters (fft_size)");
    int fft_size = 0; sscanf(argv[2], "%d", &fft_size);
    if (!incoming_block_size()) return -2;
    announce_block(fft_size);
    const size_t half = (size_t)(fft_size / 2);
    float *lower = must_alloc(sizeof(float) * half), *upper = must_alloc(sizeof(float) * half);
    for (;;) {
        if (feof(stdin)) return 0;
        fread(lower, sizeof(float), half, stdin);
        fread(upper, sizeof(float), half, stdin);
        fwrite(upper, sizeof(float), half, stdout);
        fwrite(lower, sizeof(float), half, stdout);
        end_of_block();
    }
}

static int cmd_compress_fft_adpcm_f_u8(int argc, char **argv)              /* csdr.c:1739-1767 */
{
    enum { PAD = 10 };                                                   /* the encoder needs a few values to settle: the line starts with ten copies of its first */
    if (argc <= 2) return complain("need required parameters (fft_size)");
    int fft_size = 0; sscanf(argv[2], "%d", &fft_size);
    const int line = fft_size + PAD;
    if (!incoming_block_size()) return -2;                               /* consumes a preamble if there is one (:1751) */
    announce_block(line);
    float *in = must_alloc(sizeof(float) * (size_t)line);
    short *scaled = must_alloc(sizeof(short) * (size_t)line);
    unsigned char *out = must_alloc((size_t)line / 2 + 1);
    const ima_adpcm_state_t fresh = {0, 0};
    for (;;) {
        if (feof(stdin)) return 0;
        fread(in + PAD, sizeof(float), (size_t)fft_size, stdin);
        for (int k = 0; k < PAD; k++) in[k] = in[PAD];
        for (int k = 0; k < line; k++) scaled[k] = (short)(in[k] * 100);   /* dB -> centi-dB, C conversion like the reference's */
        encode_ima_adpcm_i16_u8(scaled, out, line, fresh);               /* every line starts from the initial state (:1764) */
        fwrite(out, 1, (size_t)line / 2, stdout);
        end_of_block();
    }
}

static int cmd_encode_ima_adpcm(int argc, char **argv)                     /* csdr.c:1891-1904 */
{
    (void)argc; (void)argv;
    if (!open_block()) return -2;
    announce_block(block / 2);
    short *in = must_alloc(sizeof(short) * (size_t)block);
    unsigned char *out = must_alloc((size_t)block / 2 + 1);
    ima_adpcm_state_t st = {0, 0};
    for (;;) {
        if (feof(stdin)) return 0;
        fread(in, sizeof(short), (size_t)block, stdin);
        st = encode_ima_adpcm_i16_u8(in, out, block, st);
        fwrite(out, 1, (size_t)block / 2, stdout);
        end_of_block();
    }
}

static int cmd_limit_ff(int argc, char **argv)                              /* csdr.c:673-686 */
{
    float max_amplitude = 1.0f; if (argc >= 3) sscanf(argv[2], "%g", &max_amplitude);
    if (!announce_block(open_block())) return -2;
    float *in = must_alloc(sizeof(float) * (size_t)block), *out = must_alloc(sizeof(float) * (size_t)block);
    for (;;) {
        if (feof(stdin)) return 0;
        fread(in, sizeof(float), (size_t)block, stdin);
        limit_ff(in, out, block, max_amplitude);
        fwrite(out, sizeof(float), (size_t)block, stdout);
        end_of_block();
    }
}

static int cmd_deemphasis_wfm_ff(int argc, char **argv)                     /* csdr.c:1014-1032 */
{
    if (argc <= 3) return complain("need required parameters (sample rate, tau)");
    if (!announce_block(open_block())) return -2;
    int sample_rate = 0; sscanf(argv[2], "%d", &sample_rate);
    float tau = 0; sscanf(argv[3], "%g", &tau);
    who(); fprintf(stderr, "tau = %g, sample_rate = %d\n", tau, sample_rate);
    float *in = must_alloc(sizeof(float) * (size_t)block), *out = must_alloc(sizeof(float) * (size_t)block);
    float last = 0;
    for (;;) {
        if (feof(stdin)) return 0;
        fread(in, sizeof(float), (size_t)block, stdin);
        last = deemphasis_wfm_ff(in, out, block, tau, sample_rate, last);
        fwrite(out, sizeof(float), (size_t)block, stdout);
        end_of_block();
    }
}

static int cmd_deemphasis_nfm_ff(int argc, char **argv)                     /* csdr.c:1068-1087 */
{
    if (argc <= 2) return complain("need required parameter (sample rate)");
    int sample_rate = 0; sscanf(argv[2], "%d", &sample_rate);
    if (!announce_block(open_block())) return -2;
    /* The reference filters its still-unread buffer once before the first read (processed starts at 0, :1075-1079), so the stream
     * is effectively prefixed by one block of whatever malloc returned -- zeros in practice, zeros by construction here. */
    float *in = must_alloc(sizeof(float) * (size_t)block), *out = must_alloc(sizeof(float) * (size_t)block);
    int produced = 0;
    for (;;) {
        if (feof(stdin)) return 0;
        fread(in + block - produced, sizeof(float), (size_t)produced, stdin);
        produced = deemphasis_nfm_ff(in, out, block, sample_rate);
        if (!produced) return complain("deemphasis_nfm_ff: invalid sample rate (this function works only with specific sample rates).");
        memmove(in, in + produced, sizeof(float) * (size_t)(block - produced));
        fwrite(out, sizeof(float), (size_t)produced, stdout);
        end_of_block();
    }
}

static int cmd_fractional_decimator_ff(int argc, char **argv)
{
    if (argc <= 2) return complain("need required parameters (rate)");
    float rate = 0; sscanf(argv[2], "%g", &rate);
    int points = 12; if (argc >= 4) sscanf(argv[3], "%d", &points);
    if (points & 1) return complain("num_poly_points should be even");
    if (points < 2) return complain("num_poly_points should be >= 2");
    int prefilter = 0; float transition_bw = 0.03f; window_t window = WINDOW_DEFAULT;
    if (argc >= 5) {
        if (!strcmp(argv[4], "--prefilter")) { who(); fprintf(stderr, "using prefilter with default values\n"); prefilter = 1; }
        else { sscanf(argv[4], "%g", &transition_bw); if (argc >= 6) window = firdes_get_window_from_string(argv[5]); }
    }
    who(); fprintf(stderr, "use_prefilter = %d, num_poly_points = %d, transition_bw = %g, window = %s\n", prefilter, points, transition_bw,
                   firdes_get_string_from_window(window));
    if (!open_block()) return -2;
    announce_block((int)(block / rate));
    float *in = must_alloc(sizeof(float) * (size_t)block), *out = must_alloc(sizeof(float) * (size_t)block);
    if (rate == 1) {                                                    /* pass-through special case (:1498, clone_) */
        for (;;) { fread(in, 1, (size_t)block, stdin); fwrite(in, 1, (size_t)block, stdout); end_of_block(); if (feof(stdin)) return 0; }
    }
    int taps_length = 0; float *taps = NULL;
    if (prefilter) {
        taps_length = firdes_filter_len(transition_bw);
        who(); fprintf(stderr, "taps_length = %d\n", taps_length);
        taps = must_alloc(sizeof(float) * (size_t)taps_length);
        firdes_lowpass_f(taps, taps_length, 0.5f / (rate - transition_bw), window);
    } else { who(); fprintf(stderr, "not using taps\n"); }
    fractional_decimator_ff_t d = fractional_decimator_ff_init(rate, points, taps, taps_length);
    for (;;) {
        if (feof(stdin)) return 0;
        if (d.input_processed == 0) d.input_processed = block;
        else memcpy(in, in + d.input_processed, sizeof(float) * (size_t)(block - d.input_processed));
        fread(in + (block - d.input_processed), sizeof(float), (size_t)d.input_processed, stdin);
        fractional_decimator_ff(in, out, block, &d);
        fwrite(out, sizeof(float), (size_t)d.output_size, stdout);
        end_of_block();
    }
}

static int cmd_fastagc_ff(int argc, char **argv)
{
    static fastagc_ff_t agc;                                            /* zero-initialised like the reference's .bss copy */
    agc.input_size = 1024; if (argc >= 3) sscanf(argv[2], "%d", &agc.input_size);
    incoming_block_size();                                              /* consumes the preamble if there is one (:1385) */
    announce_block(agc.input_size);
    agc.reference = 1.0f; if (argc >= 4) sscanf(argv[3], "%g", &agc.reference);
    agc.buffer_1 = must_alloc(sizeof(float) * (size_t)agc.input_size);
    agc.buffer_2 = must_alloc(sizeof(float) * (size_t)agc.input_size);
    agc.buffer_input = must_alloc(sizeof(float) * (size_t)agc.input_size);
    float *out = must_alloc(sizeof(float) * (size_t)agc.input_size);
    for (;;) {
        if (feof(stdin)) return 0;
        fread(agc.buffer_input, sizeof(float), (size_t)agc.input_size, stdin);
        fastagc_ff(&agc, out);
        fwrite(out, sizeof(float), (size_t)agc.input_size, stdout);
        end_of_block();
    }
}

static int cmd_bandpass_fir_fft_cc(int argc, char **argv)
{
    float low_cut = 0, high_cut = 0, transition_bw = 0; window_t window = WINDOW_DEFAULT;
    int ctl = open_control(argc, argv);
    if (ctl) {
        while (!poll_control(ctl, "%g %g\n", &low_cut, &high_cut)) usleep(10000);
        if (argc <= 4) return complain("need more required parameters (transition_bw)");
    } else {
        if (argc <= 4) return complain("need required parameters (low_cut, high_cut, transition_bw)");
        sscanf(argv[2], "%g", &low_cut); sscanf(argv[3], "%g", &high_cut);
    }
    sscanf(argv[4], "%g", &transition_bw);
    if (argc >= 6) window = firdes_get_window_from_string(argv[5]);
    else { who(); fprintf(stderr, "window = %s\n", firdes_get_string_from_window(window)); }
    int taps_length = firdes_filter_len(transition_bw);
    int fft_size = next_pow2(taps_length);
    if (fft_size - taps_length < 200) fft_size <<= 1;
    int input_size = fft_size - taps_length + 1, overlap = taps_length - 1;
    who(); fprintf(stderr, "(fft_size = %d) = (taps_length = %d) + (input_size = %d) - 1\n(overlap_length = %d) = taps_length - 1\n",
                   fft_size, taps_length, input_size, overlap);
    if (fft_size <= 2) return complain("FFT size error.");
    if (!announce_block(incoming_block_size())) return -2;
    complexf *taps = must_alloc(sizeof(complexf) * (size_t)fft_size), *taps_fft = must_alloc(sizeof(complexf) * (size_t)fft_size);
    FFT_PLAN_T *plan_taps = make_fft_c2c(fft_size, taps, taps_fft, 1, 0);
    complexf *in = fft_malloc(sizeof(complexf) * (size_t)fft_size), *spec = fft_malloc(sizeof(complexf) * (size_t)fft_size);
    complexf *prod = fft_malloc(sizeof(complexf) * (size_t)fft_size);
    complexf *res[2] = {fft_malloc(sizeof(complexf) * (size_t)fft_size), fft_malloc(sizeof(complexf) * (size_t)fft_size)};
    if (!plan_taps || !in || !spec || !prod || !res[0] || !res[1]) return complain("FFT size error.");
    FFT_PLAN_T *fwd = make_fft_c2c(fft_size, in, spec, 1, 1);
    FFT_PLAN_T *inv[2] = {make_fft_c2c(fft_size, prod, res[0], 0, 1), make_fft_c2c(fft_size, prod, res[1], 0, 1)};
    memset(res[1], 0, sizeof(complexf) * (size_t)fft_size);
    memset(in, 0, sizeof(complexf) * (size_t)fft_size);
    for (;;) {
        who(); fprintf(stderr, "filter initialized, low_cut = %g, high_cut = %g\n", low_cut, high_cut);
        firdes_bandpass_c(taps, taps_length, low_cut, high_cut, window);
        fft_execute(plan_taps);
        for (int odd = 0;; odd = !odd) {
            if (feof(stdin)) return 0;
            fread(in, sizeof(complexf), (size_t)input_size, stdin);
            apply_fir_fft_cc(fwd, inv[odd], taps_fft, res[!odd] + input_size, overlap);
            fwrite(res[odd], sizeof(complexf), (size_t)input_size, stdout);
            if (poll_control(ctl, "%g %g\n", &low_cut, &high_cut)) break;
            end_of_block();
        }
    }
}

static int cmd_fastddc_fwd_cc(int argc, char **argv)
{
    if (argc <= 2) return complain("need required parameter (decimation)");
    int decimation = 0; sscanf(argv[2], "%d", &decimation);
    float transition_bw = 0.05f; if (argc > 3) sscanf(argv[3], "%g", &transition_bw);
    window_t window = WINDOW_DEFAULT;
    if (argc > 4) window = firdes_get_window_from_string(argv[4]);
    else { who(); fprintf(stderr, "window = %s\n", firdes_get_string_from_window(window)); }
    fastddc_t ddc;
    if (fastddc_init(&ddc, transition_bw, decimation, 0)) { complain("error in fastddc_init()"); return 1; }
    fastddc_print(&ddc, "fastddc_fwd_cc");
    if (!open_block()) return -2;
    announce_block(ddc.fft_size);
    complexf *in = fft_malloc(sizeof(complexf) * (size_t)ddc.fft_size), *out = fft_malloc(sizeof(complexf) * (size_t)ddc.fft_size);
    memset(in, 0, sizeof(complexf) * (size_t)ddc.fft_size);
    who(); fprintf(stderr, "benchmarking FFT...");
    FFT_PLAN_T *plan = make_fft_c2c(ddc.fft_size, in, out, 1, 1);
    fprintf(stderr, " done\n");
    if (!plan) return complain("FFT size error.");
    for (;;) {
        if (feof(stdin)) return 0;
        memmove(in, in + ddc.input_size, sizeof(complexf) * (size_t)ddc.overlap_length);      /* overlap-save (:2292) */
        fread(in + ddc.overlap_length, sizeof(complexf), (size_t)ddc.input_size, stdin);
        fft_execute(plan);                                                                       /* no window (:2295) */
        fwrite(out, sizeof(complexf), (size_t)ddc.fft_size, stdout);
        end_of_block();
    }
}

static int cmd_fastddc_inv_cc(int argc, char **argv)
{
    float shift_rate = 0; int plus = 0;
    int ctl = open_control(argc, argv);
    if (ctl) { while (!poll_control(ctl, "%g\n", &shift_rate)) usleep(10000); plus = 1; }
    else { if (argc <= 2) return complain("need required parameter (rate)"); sscanf(argv[2], "%g", &shift_rate); }
    if (argc <= 3 + plus) return complain("need required parameter (decimation)");
    int decimation = 0; sscanf(argv[3 + plus], "%d", &decimation);
    float transition_bw = 0.05f; if (argc > 4 + plus) sscanf(argv[4 + plus], "%g", &transition_bw);
    window_t window = WINDOW_DEFAULT;
    if (argc > 5 + plus) window = firdes_get_window_from_string(argv[5 + plus]);
    else { who(); fprintf(stderr, "window = %s\n", firdes_get_string_from_window(window)); }
    for (;;) {
        fastddc_t ddc;
        if (fastddc_init(&ddc, transition_bw, decimation, shift_rate)) { complain("error in fastddc_init()"); return 1; }
        fastddc_print(&ddc, "fastddc_inv_cc");
        if (!open_block()) return -2;
        announce_block(ddc.post_input_size / ddc.post_decimation);
        complexf *taps = must_alloc(sizeof(complexf) * (size_t)ddc.fft_size), *taps_fft = must_alloc(sizeof(complexf) * (size_t)ddc.fft_size);
        FFT_PLAN_T *plan_taps = make_fft_c2c(ddc.fft_size, taps, taps_fft, 1, 0);
        if (!plan_taps) return complain("FFT size error.");
        float half_bw = 0.5 / decimation;
        who(); fprintf(stderr, "preparing a bandpass filter of [%g, %g] cutoff rates. Real transition bandwidth is: %g\n",
                       (-shift_rate) - half_bw, (-shift_rate) + half_bw, 4.0 / ddc.taps_length);
        firdes_bandpass_c(taps, ddc.taps_length, (-shift_rate) - half_bw, (-shift_rate) + half_bw, window);
        fft_execute(plan_taps);
        fft_swap_sides(taps_fft, ddc.fft_size);
        complexf *inv_in = fft_malloc(sizeof(complexf) * (size_t)ddc.fft_inv_size), *inv_out = fft_malloc(sizeof(complexf) * (size_t)ddc.fft_inv_size);
        who(); fprintf(stderr, "benchmarking FFT...");
        FFT_PLAN_T *plan_inverse = make_fft_c2c(ddc.fft_inv_size, inv_in, inv_out, 0, 1);
        fprintf(stderr, " done\n");
        complexf *in = fft_malloc(sizeof(complexf) * (size_t)ddc.fft_size), *out = fft_malloc(sizeof(complexf) * (size_t)ddc.post_input_size);
        decimating_shift_addition_status_t st; memset(&st, 0, sizeof st);
        for (;;) {
            if (feof(stdin)) return 0;
            fread(in, sizeof(complexf), (size_t)ddc.fft_size, stdin);
            st = fastddc_inv_cc(in, out, &ddc, plan_inverse, taps_fft, st);
            fwrite(out, sizeof(complexf), (size_t)st.output_size, stdout);
            end_of_block();
            if (poll_control(ctl, "%g\n", &shift_rate)) break;
        }
        free(taps); free(taps_fft); fft_destroy(plan_taps); fft_destroy(plan_inverse);
        fft_free(inv_in); fft_free(inv_out); fft_free(in); fft_free(out);
    }
}

/* ---- dispatch ------------------------------------------------------------------------------------ */
static const struct { const char *name; int (*run)(int, char **); const char *syntax; } kCommands[] = {
    {"convert_u8_f", cmd_convert_u8_f, "convert_u8_f"},
    {"convert_s16_f", cmd_convert_s16_f, "convert_s16_f"},
    {"convert_i16_f", cmd_convert_s16_f, "convert_i16_f"},
    {"convert_f_s16", cmd_convert_f_s16, "convert_f_s16"},
    {"convert_f_i16", cmd_convert_f_s16, "convert_f_i16"},
    {"shift_addition_cc", cmd_shift_addition_cc, "shift_addition_cc <rate> | --fifo <fifo_path> | --fd <fd>"},
    {"fir_decimate_cc", cmd_fir_decimate_cc, "fir_decimate_cc <decimation_factor> [transition_bw [window]]"},
    {"fmdemod_quadri_cf", cmd_fmdemod_quadri_cf, "fmdemod_quadri_cf"},
    {"fractional_decimator_ff", cmd_fractional_decimator_ff, "fractional_decimator_ff <decimation_rate> [num_poly_points ( [transition_bw [window]] | --prefilter )]"},
    {"fastagc_ff", cmd_fastagc_ff, "fastagc_ff [block_size [reference]]"},
    {"limit_ff", cmd_limit_ff, "limit_ff [max_amplitude]"},
    {"fft_exchange_sides_ff", cmd_fft_exchange_sides_ff, "fft_exchange_sides_ff <fft_size>"},
    {"compress_fft_adpcm_f_u8", cmd_compress_fft_adpcm_f_u8, "compress_fft_adpcm_f_u8 <fft_size>"},
    {"encode_ima_adpcm_i16_u8", cmd_encode_ima_adpcm, "encode_ima_adpcm_i16_u8"},
    {"encode_ima_adpcm_s16_u8", cmd_encode_ima_adpcm, "encode_ima_adpcm_s16_u8"},
    {"shift_unroll_cc", cmd_shift_unroll_cc, "shift_unroll_cc <rate> | --fifo <fifo_path> | --fd <fd>"},
    {"shift_math_cc", cmd_shift_math_cc, "shift_math_cc <rate>"},
    {"shift_table_cc", cmd_shift_table_cc, "shift_table_cc <rate> [table_size]"},
    {"shift_addfast_cc", cmd_shift_addfast_cc, "shift_addfast_cc <rate> | --fifo <fifo_path> | --fd <fd>"},
    {"decimating_shift_addition_cc", cmd_decimating_shift_addition_cc, "decimating_shift_addition_cc <rate> [decimation]"},
    {"fft_cc", cmd_fft_cc, "fft_cc <fft_size> <out_of_every_n_samples> [window]"},
    {"logpower_cf", cmd_logpower_cf, "logpower_cf [add_db]"},
    {"logaveragepower_cf", cmd_logaveragepower_cf, "logaveragepower_cf <add_db> <fft_size> <avgnumber>"},
    {"deemphasis_wfm_ff", cmd_deemphasis_wfm_ff, "deemphasis_wfm_ff <sample_rate> <tau>"},
    {"deemphasis_nfm_ff", cmd_deemphasis_nfm_ff, "deemphasis_nfm_ff <one_of_the_predefined_sample_rates>"},
    {"bandpass_fir_fft_cc", cmd_bandpass_fir_fft_cc, "bandpass_fir_fft_cc <low_cut> <high_cut> <transition_bw> [window] | --fifo <fifo_path> <transition_bw> [window]"},
    {"fastddc_fwd_cc", cmd_fastddc_fwd_cc, "fastddc_fwd_cc <decimation> [transition_bw [window]]"},
    {"fastddc_inv_cc", cmd_fastddc_inv_cc, "fastddc_inv_cc <shift_rate> <decimation> [transition_bw [window]] | --fifo <fifo_path> ... | --fd <fd> ..."},
};

static int usage(void)
{
    fprintf(stderr, "csdr (B200 hot-path build) - DSP blocks on stdin/stdout, computed by libcsdr_b200 on a CUDA device\nusage:\n");
    for (size_t k = 0; k < sizeof kCommands / sizeof kCommands[0]; k++) fprintf(stderr, "    csdr %s\n", kCommands[k].syntax);
    fprintf(stderr, "commands of the reference CLI outside this list are not part of this build\n");
    return -1;
}

int main(int argc, char **argv)
{
    read_environment();
    G.argc = argc; G.argv = argv;
    if (argc <= 1 || !strcmp(argv[1], "--help")) return usage();
    fcntl(STDIN_FILENO, F_SETPIPE_SZ, 65536 * 32);
    fcntl(STDOUT_FILENO, F_SETPIPE_SZ, 65536 * 32);
    for (size_t k = 0; k < sizeof kCommands / sizeof kCommands[0]; k++)
        if (!strcmp(argv[1], kCommands[k].name)) return kCommands[k].run(argc, argv);
    return complain("function name given in argument 1 does not exist (in this hot-path build). Possible causes: you have mistyped the commmand name, "
                    "or the command belongs to the reference CLI only.");
}
