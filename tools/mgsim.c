/* mgsim -- deterministic synthetic workloads for the mapping benchmarks.
 *
 *   mgsim graph  -l 5000000 -n 8 -s 7 -o PREFIX     random backbone + SV haplotypes -> PREFIX.gfa (rGFA), PREFIX.hap.fa
 *   mgsim walk   -g X.gfa -w ">a>b<c" [-w ...] -o out.hap.fa   spell haplotypes out of a GFA by walking it
 *   mgsim reads  -i hap.fa -n 100000 -l 15000 -e ont|hifi -s 5 [-c] -o reads.fa   sample error-carrying reads
 *
 * PRNG: splitmix64 only (explicit seeds; no libc rand), so that every box generates byte-identical inputs.
 * Error models (iid per base): ont = 4% substitution, 3% deletion, 3% 1-bp insertion; hifi = 0.2/0.15/0.15 %.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

static uint64_t sm_state;
static inline uint64_t sm_next(void)
{
	uint64_t z = (sm_state += 0x9E3779B97F4A7C15ULL);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}
static inline double sm_unif(void) { return (sm_next() >> 11) * (1.0 / 9007199254740992.0); }
static inline uint64_t sm_below(uint64_t n) { return (uint64_t)(sm_unif() * (double)n); }

static const char NT[5] = "ACGT";
static inline char comp(char c) { return c == 'A'? 'T' : c == 'C'? 'G' : c == 'G'? 'C' : c == 'T'? 'A' : c; }

typedef struct { size_t l, m; char *s; } str_t;
static void str_push(str_t *s, const char *p, size_t n)
{
	if (s->l + n + 1 > s->m) {
		s->m = (s->l + n + 1) * 3 / 2 + 64;
		s->s = (char*)realloc(s->s, s->m);
	}
	memcpy(s->s + s->l, p, n);
	s->l += n, s->s[s->l] = 0;
}
static void str_push_rc(str_t *s, const char *p, size_t n)
{
	size_t i;
	str_push(s, p, n);
	for (i = 0; i < n; ++i) s->s[s->l - n + i] = comp(p[n - 1 - i]);
}

static void write_fa(FILE *fp, const char *name, const char *seq, size_t len)
{
	fprintf(fp, ">%s\n", name);
	fwrite(seq, 1, len, fp);
	fputc('\n', fp);
}

/******************** graph ********************/

typedef struct {
	int64_t s, e;      /* backbone interval [s,e); e==s for insertions */
	int type;          /* 0 ins, 1 del, 2 inv, 3 replacement */
	int alt_len;
	char *alt;
	uint32_t haps;     /* bit h set: haplotype h carries the event */
	int seg_alt;       /* segment id of the alt sequence (ins/replacement) */
} ev_t;

static int cmd_graph(int argc, char **argv)
{
	int64_t L = 1000000, i, pos;
	int n_hap = 3, k, n_ev = 0, m_ev = 0;
	uint64_t seed = 7;
	const char *prefix = "sim";
	double rate = 1.0 / 4000.0, snp = 0.001;
	static const int sv_len[6] = { 60, 120, 300, 800, 2000, 5000 };
	char *bb, fn[1024];
	ev_t *ev = 0;
	FILE *fg, *fh;
	for (k = 2; k < argc; ++k) {
		if (!strcmp(argv[k], "-l")) L = atoll(argv[++k]);
		else if (!strcmp(argv[k], "-n")) n_hap = atoi(argv[++k]);
		else if (!strcmp(argv[k], "-s")) seed = strtoull(argv[++k], 0, 10);
		else if (!strcmp(argv[k], "-o")) prefix = argv[++k];
	}
	if (n_hap > 31) n_hap = 31;
	sm_state = seed;
	bb = (char*)malloc(L + 1);
	for (i = 0; i < L; ++i) bb[i] = NT[sm_next() >> 62];
	bb[L] = 0;
	/* events: walk along the backbone; gaps between events are geometric-ish with mean chosen such that each
	 * haplotype sees ~rate events per base when every event is carried by on average 2 haplotypes */
	{
		double per_ev_haps = n_hap >= 2? 2.0 : 1.0;
		double ev_rate = rate * n_hap / per_ev_haps; /* events per backbone base */
		double mean_gap = 1.0 / ev_rate;
		pos = 500;
		while (pos < L - 6000) {
			ev_t e;
			double u = sm_unif();
			int len = sv_len[sm_below(6)], h;
			memset(&e, 0, sizeof(e));
			e.type = u < 0.4? 0 : u < 0.8? 1 : u < 0.9? 2 : 3;
			e.s = pos, e.e = e.type == 0? pos : pos + len;
			if (e.type == 0 || e.type == 3) {
				int64_t j;
				e.alt_len = e.type == 0? len : sv_len[sm_below(6)];
				e.alt = (char*)malloc(e.alt_len + 1);
				for (j = 0; j < e.alt_len; ++j) e.alt[j] = NT[sm_next() >> 62];
				e.alt[e.alt_len] = 0;
			}
			for (h = 0; h < n_hap; ++h)
				if (sm_unif() < per_ev_haps / n_hap) e.haps |= 1u << h;
			if (e.haps == 0) e.haps = 1u << sm_below(n_hap);
			if (n_ev == m_ev) { m_ev = m_ev? m_ev * 2 : 1024; ev = (ev_t*)realloc(ev, m_ev * sizeof(ev_t)); }
			ev[n_ev++] = e;
			/* next event: at least 150 bp of untouched backbone after this one */
			pos = e.e + 150 + (int64_t)(-mean_gap * 0.9 * __builtin_log(1.0 - sm_unif()));
		}
	}
	/* segments: backbone pieces between consecutive breakpoints; event k of type del/inv/rep owns piece [s,e) */
	snprintf(fn, sizeof(fn), "%s.gfa", prefix);
	fg = fopen(fn, "w");
	snprintf(fn, sizeof(fn), "%s.hap.fa", prefix);
	fh = fopen(fn, "w");
	if (fg == 0 || fh == 0) { fprintf(stderr, "mgsim: cannot write %s.*\n", prefix); return 1; }
	{
		/* piece list: boundaries = 0, every ev.s, every ev.e, L */
		int64_t *bd = (int64_t*)malloc((2 * (size_t)n_ev + 2) * sizeof(int64_t));
		int n_bd = 0, n_piece, p, sid = 0;
		int *piece_of_ev_start = (int*)malloc((size_t)n_ev * sizeof(int)); /* piece index that begins at ev.s */
		bd[n_bd++] = 0;
		for (k = 0; k < n_ev; ++k) {
			if (bd[n_bd - 1] != ev[k].s) bd[n_bd++] = ev[k].s;
			piece_of_ev_start[k] = n_bd - 1;
			if (ev[k].e != ev[k].s) bd[n_bd++] = ev[k].e;
		}
		bd[n_bd++] = L;
		n_piece = n_bd - 1;
		for (p = 0; p < n_piece; ++p) { /* backbone S-lines: segment p is named s{p+1} */
			fprintf(fg, "S\ts%d\t", p + 1);
			fwrite(bb + bd[p], 1, bd[p + 1] - bd[p], fg);
			fprintf(fg, "\tLN:i:%ld\tSN:Z:chr1\tSO:i:%ld\tSR:i:0\n", (long)(bd[p + 1] - bd[p]), (long)bd[p]);
		}
		sid = n_piece;
		for (k = 0; k < n_ev; ++k) {
			if (ev[k].alt) {
				int h, rank = 1;
				for (h = 0; h < n_hap; ++h) if (ev[k].haps >> h & 1) { rank = h + 1; break; }
				ev[k].seg_alt = ++sid;
				fprintf(fg, "S\ts%d\t%s\tLN:i:%d\tSN:Z:hap%d_e%d\tSO:i:0\tSR:i:%d\n", sid, ev[k].alt, ev[k].alt_len, rank, k, rank);
			}
		}
		for (p = 0; p + 1 < n_piece; ++p)
			fprintf(fg, "L\ts%d\t+\ts%d\t+\t0M\tSR:i:0\n", p + 1, p + 2);
		for (k = 0; k < n_ev; ++k) {
			int ps = piece_of_ev_start[k]; /* piece starting at ev.s */
			int left = ps - 1;             /* piece ending at ev.s (always exists: events start >= 500) */
			int h, rank = 1;
			for (h = 0; h < n_hap; ++h) if (ev[k].haps >> h & 1) { rank = h + 1; break; }
			if (ev[k].type == 0) { /* insertion between left and ps */
				fprintf(fg, "L\ts%d\t+\ts%d\t+\t0M\tSR:i:%d\n", left + 1, ev[k].seg_alt, rank);
				fprintf(fg, "L\ts%d\t+\ts%d\t+\t0M\tSR:i:%d\n", ev[k].seg_alt, ps + 1, rank);
			} else if (ev[k].type == 1) { /* deletion: left -> piece after ps */
				fprintf(fg, "L\ts%d\t+\ts%d\t+\t0M\tSR:i:%d\n", left + 1, ps + 2, rank);
			} else if (ev[k].type == 2) { /* inversion of piece ps */
				fprintf(fg, "L\ts%d\t+\ts%d\t-\t0M\tSR:i:%d\n", left + 1, ps + 1, rank);
				fprintf(fg, "L\ts%d\t-\ts%d\t+\t0M\tSR:i:%d\n", ps + 1, ps + 2, rank);
			} else { /* replacement of piece ps by alt */
				fprintf(fg, "L\ts%d\t+\ts%d\t+\t0M\tSR:i:%d\n", left + 1, ev[k].seg_alt, rank);
				fprintf(fg, "L\ts%d\t+\ts%d\t+\t0M\tSR:i:%d\n", ev[k].seg_alt, ps + 2, rank);
			}
		}
		free(bd); free(piece_of_ev_start);
	}
	/* haplotypes: backbone itself + n_hap mutated copies */
	write_fa(fh, "hap0", bb, L);
	{
		int h;
		for (h = 0; h < n_hap; ++h) {
			str_t s = {0,0,0};
			char name[64];
			int64_t x = 0;
			size_t j;
			for (k = 0; k < n_ev; ++k) {
				if (!(ev[k].haps >> h & 1)) continue;
				str_push(&s, bb + x, ev[k].s - x);
				if (ev[k].type == 0) str_push(&s, ev[k].alt, ev[k].alt_len);
				else if (ev[k].type == 2) str_push_rc(&s, bb + ev[k].s, ev[k].e - ev[k].s);
				else if (ev[k].type == 3) str_push(&s, ev[k].alt, ev[k].alt_len);
				x = ev[k].e;
			}
			str_push(&s, bb + x, L - x);
			for (j = 0; j < s.l; ++j) /* SNPs */
				if (sm_unif() < snp) {
					int c = (int)(sm_next() >> 62);
					if (NT[c] == s.s[j]) c = (c + 1) & 3;
					s.s[j] = NT[c];
				}
			snprintf(name, sizeof(name), "hap%d", h + 1);
			write_fa(fh, name, s.s, s.l);
			free(s.s);
		}
	}
	fclose(fg); fclose(fh);
	fprintf(stderr, "[mgsim] backbone %ld bp, %d events, %d haplotypes -> %s.gfa %s.hap.fa\n", (long)L, n_ev, n_hap, prefix, prefix);
	return 0;
}

/******************** FASTA reader (plain, multi-line ok) ********************/

typedef struct { char *name; char *seq; int64_t len; } fa1_t;

static fa1_t *read_fa(const char *fn, int *n_)
{
	FILE *fp = fopen(fn, "r");
	fa1_t *a = 0;
	int n = 0, m = 0;
	char *line = 0;
	size_t cap = 0;
	ssize_t l;
	str_t cur = {0,0,0};
	if (fp == 0) { *n_ = 0; return 0; }
	while ((l = getline(&line, &cap, fp)) >= 0) {
		while (l > 0 && (line[l-1] == '\n' || line[l-1] == '\r')) line[--l] = 0;
		if (l > 0 && line[0] == '>') {
			char *p;
			if (n > 0) a[n-1].seq = cur.s, a[n-1].len = cur.l, memset(&cur, 0, sizeof(cur));
			if (n == m) { m = m? m * 2 : 16; a = (fa1_t*)realloc(a, m * sizeof(fa1_t)); }
			for (p = line + 1; *p && *p != ' ' && *p != '\t'; ++p) {}
			*p = 0;
			a[n].name = strdup(line + 1), a[n].seq = 0, a[n].len = 0;
			++n;
		} else if (n > 0) str_push(&cur, line, l);
	}
	if (n > 0) a[n-1].seq = cur.s, a[n-1].len = cur.l;
	free(line);
	fclose(fp);
	*n_ = n;
	return a;
}

/******************** walk ********************/

static int cmd_walk(int argc, char **argv)
{
	const char *gfa = 0, *out = 0, *walks[64];
	int n_w = 0, k, n_seg = 0, m_seg = 0, w;
	char **name = 0, **seq = 0, *line = 0;
	size_t cap = 0;
	ssize_t l;
	FILE *fp, *fo;
	for (k = 2; k < argc; ++k) {
		if (!strcmp(argv[k], "-g")) gfa = argv[++k];
		else if (!strcmp(argv[k], "-o")) out = argv[++k];
		else if (!strcmp(argv[k], "-w") && n_w < 64) walks[n_w++] = argv[++k];
	}
	if (gfa == 0 || out == 0 || n_w == 0) return 1;
	fp = fopen(gfa, "r");
	if (fp == 0) { fprintf(stderr, "mgsim: cannot open %s\n", gfa); return 1; }
	while ((l = getline(&line, &cap, fp)) >= 0) {
		char *p, *q;
		if (l < 3 || line[0] != 'S' || line[1] != '\t') continue;
		p = line + 2;
		for (q = p; *q && *q != '\t'; ++q) {}
		*q++ = 0;
		if (n_seg == m_seg) { m_seg = m_seg? m_seg * 2 : 64; name = (char**)realloc(name, m_seg * sizeof(char*)); seq = (char**)realloc(seq, m_seg * sizeof(char*)); }
		name[n_seg] = strdup(p);
		for (p = q; *q && *q != '\t' && *q != '\n'; ++q) {}
		*q = 0;
		seq[n_seg++] = strdup(p);
	}
	fclose(fp);
	fo = fopen(out, "w");
	for (w = 0; w < n_w; ++w) {
		const char *p = walks[w];
		str_t s = {0,0,0};
		char nm[32];
		while (*p == '>' || *p == '<') {
			int rev = *p == '<', i;
			const char *q = ++p;
			while (*p && *p != '>' && *p != '<') ++p;
			for (i = 0; i < n_seg; ++i)
				if (strlen(name[i]) == (size_t)(p - q) && strncmp(name[i], q, p - q) == 0) break;
			if (i == n_seg) { fprintf(stderr, "mgsim: unknown segment in walk %s\n", walks[w]); return 1; }
			if (rev) str_push_rc(&s, seq[i], strlen(seq[i]));
			else str_push(&s, seq[i], strlen(seq[i]));
		}
		snprintf(nm, sizeof(nm), "walk%d", w);
		write_fa(fo, nm, s.s, s.l);
		free(s.s);
	}
	fclose(fo);
	return 0;
}

/******************** reads ********************/

static int cmd_reads(int argc, char **argv)
{
	const char *in = 0, *out = 0, *err = "ont";
	int64_t n_reads = 1000, rlen = 10000, r;
	uint64_t seed = 11;
	int k, n_hap, circ = 0;
	double p_sub, p_del, p_ins;
	fa1_t *hap;
	FILE *fo;
	char *buf;
	int64_t tot = 0;
	for (k = 2; k < argc; ++k) {
		if (!strcmp(argv[k], "-i")) in = argv[++k];
		else if (!strcmp(argv[k], "-o")) out = argv[++k];
		else if (!strcmp(argv[k], "-n")) n_reads = atoll(argv[++k]);
		else if (!strcmp(argv[k], "-l")) rlen = atoll(argv[++k]);
		else if (!strcmp(argv[k], "-e")) err = argv[++k];
		else if (!strcmp(argv[k], "-s")) seed = strtoull(argv[++k], 0, 10);
		else if (!strcmp(argv[k], "-c")) circ = 1; /* circularise: sample from the sequence concatenated twice */
	}
	if (in == 0 || out == 0) return 1;
	if (!strcmp(err, "hifi")) p_sub = 0.002, p_del = 0.0015, p_ins = 0.0015;
	else if (!strcmp(err, "none")) p_sub = p_del = p_ins = 0.0;
	else p_sub = 0.04, p_del = 0.03, p_ins = 0.03;
	hap = read_fa(in, &n_hap);
	if (hap == 0 || n_hap == 0) { fprintf(stderr, "mgsim: no sequence in %s\n", in); return 1; }
	if (circ)
		for (k = 0; k < n_hap; ++k) {
			hap[k].seq = (char*)realloc(hap[k].seq, 2 * hap[k].len + 1);
			memcpy(hap[k].seq + hap[k].len, hap[k].seq, hap[k].len);
			hap[k].len *= 2, hap[k].seq[hap[k].len] = 0;
		}
	for (k = 0; k < n_hap; ++k) { /* upper-case; anything but ACGT stays as is */
		int64_t j;
		for (j = 0; j < hap[k].len; ++j) if (hap[k].seq[j] >= 'a' && hap[k].seq[j] <= 'z') hap[k].seq[j] -= 32;
	}
	sm_state = seed;
	fo = fopen(out, "w");
	buf = (char*)malloc(rlen * 2 + 16);
	for (r = 0; r < n_reads; ++r) {
		int h = (int)sm_below(n_hap), rev;
		int64_t L = hap[h].len, len = rlen < L? rlen : L, st, i, m = 0;
		const char *s;
		st = (int64_t)sm_below(L - len + 1);
		rev = (int)(sm_next() >> 63);
		s = hap[h].seq + st;
		for (i = 0; i < len; ++i) {
			char c = rev? comp(s[len - 1 - i]) : s[i];
			double u = sm_unif();
			if (u < p_sub) {
				int d = (int)(sm_next() >> 62);
				if (NT[d] == c) d = (d + 1) & 3;
				buf[m++] = NT[d];
			} else if (u < p_sub + p_del) {
				/* deleted */
			} else if (u < p_sub + p_del + p_ins) {
				buf[m++] = c;
				buf[m++] = NT[sm_next() >> 62];
			} else buf[m++] = c;
		}
		fprintf(fo, ">r%ld_%s_%ld_%c\n", (long)r, hap[h].name, (long)st, "+-"[rev]);
		fwrite(buf, 1, m, fo);
		fputc('\n', fo);
		tot += m;
	}
	fclose(fo);
	fprintf(stderr, "[mgsim] %ld reads, %ld bases -> %s\n", (long)n_reads, (long)tot, out);
	return 0;
}

int main(int argc, char **argv)
{
	if (argc < 2) {
		fprintf(stderr, "Usage: mgsim <graph|walk|reads> [options]\n");
		return 1;
	}
	if (!strcmp(argv[1], "graph")) return cmd_graph(argc, argv);
	if (!strcmp(argv[1], "walk")) return cmd_walk(argc, argv);
	if (!strcmp(argv[1], "reads")) return cmd_reads(argc, argv);
	fprintf(stderr, "mgsim: unknown command %s\n", argv[1]);
	return 1;
}
