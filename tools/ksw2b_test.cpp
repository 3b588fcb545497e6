// ksw2b-test -- command-line front end of ksw2_b200 with the options and the output format of the reference's `ksw2-test`
// (cli.c:141-260): two FASTA files (plain or gzip) or two literal sequences in, one line per pair out:
//     tname  qname  score  max  max_t  max_q  [CIGAR]
// Unlike the reference, which aligns pair after pair (cli.c:223-224), the pairs of a run are collected and aligned as ONE batch
// through the C ABI (ksw2b_align) -- the B200 wants thousands of pairs per launch.  Host code only; all alignment work is done
// by libksw2_b200.so on the GPU.  CIGAR ops are printed as M/I/D/N/=/X (the reference indexes "MID" and prints a NUL byte for N).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <zlib.h>
#include <string>
#include <vector>
#include "../include/ksw2_b200.h"

struct Rec { std::string name, seq; };

static uint8_t nt4(unsigned char c)                      // cli.c:17-34: 0..3 stay, A/C/G/T in either case -> 0..3, anything else -> 4
{
	switch (c) { case 0: case 'A': case 'a': return 0; case 1: case 'C': case 'c': return 1; case 2: case 'G': case 'g': return 2;
	             case 3: case 'T': case 't': return 3; default: return 4; }
}

// FASTA / FASTQ records from a (possibly gzip-compressed) file; the name is the first word of the header line
static bool read_all(const char *path, std::vector<Rec> &out)
{
	gzFile fp = gzopen(path, "r");
	if (!fp) return false;
	std::string line, cur;
	std::vector<char> buf(1 << 16);
	Rec r; bool have = false, fastq = false; int qual_left = 0; bool in_qual = false;
	auto flush = [&]() { if (have) out.push_back(r); r = Rec(); have = false; };
	while (gzgets(fp, buf.data(), (int)buf.size())) {
		line.assign(buf.data());
		while (!line.empty() && line.back() != '\n' && gzgets(fp, buf.data(), (int)buf.size())) line += buf.data();   // very long lines
		while (!line.empty() && (line.back() == '\n' || line.back() == '\r')) line.pop_back();
		if (in_qual) { qual_left -= (int)line.size(); if (qual_left <= 0) in_qual = false; continue; }
		if (line.empty()) continue;
		if (line[0] == '>' || (line[0] == '@' && !have) || (line[0] == '@' && fastq)) {
			flush();
			fastq = line[0] == '@';
			size_t e = 1; while (e < line.size() && line[e] != ' ' && line[e] != '\t') ++e;
			r.name = line.substr(1, e - 1); have = true;
		} else if (line[0] == '+' && fastq) { in_qual = true; qual_left = (int)r.seq.size(); }
		else if (have) r.seq += line;
	}
	flush();
	gzclose(fp);
	return true;
}

static void simple_mat(int8_t *mat, int a, int b)          // what cli.c:36-48 builds: a on the diagonal, -|b| elsewhere, 0 for the wildcard
{
	a = a < 0 ? -a : a; b = b > 0 ? -b : b;
	for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) mat[i * 5 + j] = (i == 4 || j == 4) ? 0 : (i == j ? a : b);
}

int main(int argc, char *argv[])
{
	int a = 2, b = 4, q = 4, e = 2, q2 = 13, e2 = 1, c, pair = 1, w = -1, flag = 0, rep = 1, zdrop = -1;
	const char *algo = "extd";
	char *s;
	while ((c = getopt(argc, argv, "t:w:R:rsgz:A:B:O:E:Ka")) >= 0) {
		if (c == 't') algo = optarg;
		else if (c == 'w') w = atoi(optarg);
		else if (c == 'R') rep = atoi(optarg);
		else if (c == 'z') zdrop = atoi(optarg);
		else if (c == 'r') flag |= KSW_EZ_RIGHT;
		else if (c == 's') flag |= KSW_EZ_SCORE_ONLY;
		else if (c == 'g') flag |= KSW_EZ_APPROX_MAX | KSW_EZ_APPROX_DROP;
		else if (c == 'K') {}                                            // the reference's "no kalloc" switch: nothing to switch here
		else if (c == 'A') a = atoi(optarg);
		else if (c == 'B') b = atoi(optarg);
		else if (c == 'a') pair = 0;
		else if (c == 'O') { q = q2 = (int)strtol(optarg, &s, 10); if (*s == ',') q2 = (int)strtol(s + 1, &s, 10); }
		else if (c == 'E') { e = e2 = (int)strtol(optarg, &s, 10); if (*s == ',') e2 = (int)strtol(s + 1, &s, 10); }
	}
	if (argc - optind < 2) {
		fprintf(stderr, "Usage: ksw2b-test [options] <DNA-target> <DNA-query>\n");
		fprintf(stderr, "Options:\n");
		fprintf(stderr, "  -t STR        algorithm: gg, gg2, gg2_sse, extz, extz2_sse, extd, extd2_sse, extf2_sse, exts2_sse, test [%s]\n", algo);
		fprintf(stderr, "  -R INT        repeat the batch INT times (for benchmarking) [1]\n");
		fprintf(stderr, "  -w INT        band width [inf]\n  -z INT        Z-drop [%d]\n  -r            gap right alignment\n  -s            score only\n", zdrop);
		fprintf(stderr, "  -g            approximate max / drop (KSW_EZ_APPROX_MAX|KSW_EZ_APPROX_DROP)\n");
		fprintf(stderr, "  -A INT        match score [%d]\n  -B INT        mismatch penalty [%d]\n", a, b);
		fprintf(stderr, "  -O INT[,INT]  gap open penalty [%d,%d]\n  -E INT[,INT]  gap extension penalty [%d,%d]\n  -a            all vs all\n", q, q2, e, e2);
		return 1;
	}
	// ---- which entry point, with which arguments (cli.c:66-84) ----
	int8_t mat[25];
	simple_mat(mat, a, -b);
	ksw2b_params_t P; memset(&P, 0, sizeof P);
	P.m = 5; P.mat = mat; P.q = q; P.e = e; P.q2 = q2; P.e2 = e2; P.w = w; P.zdrop = zdrop; P.end_bonus = 0; P.flag = flag;
	bool gg = false;
	if (!strcmp(algo, "gg")) { P.kind = KSW2B_GG; gg = true; }
	else if (!strcmp(algo, "gg2")) { P.kind = KSW2B_GG2; gg = true; }
	else if (!strcmp(algo, "gg2_sse")) { P.kind = KSW2B_GG2_SSE; gg = true; P.flag &= ~KSW_EZ_SCORE_ONLY; }      // cli.c:74 always passes the CIGAR pointers
	else if (!strcmp(algo, "extz")) P.kind = KSW2B_EXTZ;
	else if (!strcmp(algo, "extz2_sse")) P.kind = KSW2B_EXTZ2;
	else if (!strcmp(algo, "extd")) P.kind = KSW2B_EXTD;
	else if (!strcmp(algo, "extd2_sse")) P.kind = KSW2B_EXTD2;
	else if (!strcmp(algo, "extf2_sse")) { P.kind = KSW2B_EXTF2; P.q = mat[0]; P.q2 = mat[1]; P.flag = KSW_EZ_SCORE_ONLY; }
	else if (!strcmp(algo, "exts2_sse")) { P.kind = KSW2B_EXTS2; simple_mat(mat, 1, 2); P.q = 2; P.e = 1; P.q2 = 32; P.noncan = 4; P.junc_bonus = 0; P.flag = flag | KSW_EZ_SPLICE_FOR; }
	else if (!strcmp(algo, "test")) { P.kind = KSW2B_EXTD2; P.q = 4; P.e = 2; P.q2 = 24; P.e2 = 1; P.w = 751; P.zdrop = 400; P.flag = 8; }
	else { fprintf(stderr, "ERROR: can't find algorithm '%s'\n", algo); return 1; }
	if (gg) { P.zdrop = -1; P.flag &= KSW_EZ_SCORE_ONLY; }
	// ---- the pairs, in the reference's output order ----
	std::vector<Rec> T, Q;
	std::vector<std::pair<int, int> > pairs;      // (target index, query index)
	const bool f0 = read_all(argv[optind], T), f1 = read_all(argv[optind + 1], Q);
	if (!f0 && !f1) {                                                   // literal sequences (cli.c:216-218)
		T.push_back(Rec{"first", argv[optind]}); Q.push_back(Rec{"second", argv[optind + 1]}); pairs.push_back({0, 0});
	} else if (f0 && f1) {
		if (pair) { for (size_t i = 0; i < T.size() && i < Q.size(); ++i) pairs.push_back({(int)i, (int)i}); }
		else { for (size_t j = 0; j < Q.size(); ++j) for (size_t i = 0; i < T.size(); ++i) pairs.push_back({(int)i, (int)j}); }
	} else return 0;                                                    // one file only: the reference prints nothing
	const int64_t n = (int64_t)pairs.size();
	std::vector<int64_t> qoff((size_t)n + 1, 0), toff((size_t)n + 1, 0);
	for (int64_t i = 0; i < n; ++i) { qoff[i + 1] = qoff[i] + (int64_t)Q[pairs[i].second].seq.size(); toff[i + 1] = toff[i] + (int64_t)T[pairs[i].first].seq.size(); }
	std::vector<uint8_t> qcat((size_t)qoff[n] + 1), tcat((size_t)toff[n] + 1);
	for (int64_t i = 0; i < n; ++i) {
		const std::string &qs = Q[pairs[i].second].seq, &ts = T[pairs[i].first].seq;
		for (size_t k = 0; k < qs.size(); ++k) qcat[(size_t)qoff[i] + k] = nt4((unsigned char)qs[k]);
		for (size_t k = 0; k < ts.size(); ++k) tcat[(size_t)toff[i] + k] = nt4((unsigned char)ts[k]);
	}
	// ---- one batch on the GPU ----
	ksw2b_ctx_t *ctx = ksw2b_create(-1);
	if (!ctx) { fprintf(stderr, "ksw2b-test: %s\n", ksw2b_last_error()); return 2; }
	std::vector<ksw2b_result_t> res((size_t)n);
	const uint32_t *cig = 0;
	for (int it = 0; it < (rep > 0 ? rep : 1); ++it) {
		const int rc = ksw2b_align(ctx, &P, n, qcat.data(), qoff.data(), tcat.data(), toff.data(), 0, res.data(), &cig);
		if (rc) { fprintf(stderr, "ksw2b-test: alignment failed (%d): %s\n", rc, ksw2b_last_error()); return 2; }
	}
	for (int64_t i = 0; i < n; ++i) {                                   // print_aln, cli.c:141-152
		const ksw2b_result_t &r = res[(size_t)i];
		printf("%s\t%s\t%d\t%d\t%d\t%d", T[pairs[i].first].name.c_str(), Q[pairs[i].second].name.c_str(), r.score, r.max, r.max_t, r.max_q);
		if (r.n_cigar > 0 && cig) {
			putchar('\t');
			for (int k = 0; k < r.n_cigar; ++k) { const uint32_t x = cig[r.cigar_off + k]; printf("%u%c", x >> 4, "MIDN___=X"[x & 0xf]); }
		}
		putchar('\n');
	}
	ksw2b_destroy(ctx);
	return 0;
}
