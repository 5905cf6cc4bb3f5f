# Applied to the reference's gmap.c (read in place from /root/reference, never copied into the repo): produces
# oracle/_ref/gmap_b200.c, the ONLY host file that changes when minigraph is linked against libmgb200.so.
# (1) a helper that hands one whole mini-batch to the GPU dispatcher, (2) the kt_for call at gmap.c:99.
/^static void \*worker_pipeline\(void \*shared, int step, void \*in\)/ {
	print "extern int mg_map_batch_frag(const mg_idx_t *gi, int n_frag, const int *n_seg, const int *qlens, const char *const *seqs, const char *const *names, mg_gchains_t **gcs, const mg_mapopt_t *opt);"
	print "static void mgb_worker_batch(step_t *s) // replaces kt_for(worker_for): the mini-batch as one call (fragments of one or more segments)"
	print "{"
	print "\tint i, j, n = s->n_seq, pe_ori = s->p->opt->pe_ori, *ql = (int*)malloc(n * sizeof(int));"
	print "\tconst char **sq = (const char**)malloc(n * sizeof(char*)), **nm = (const char**)malloc(s->n_frag * sizeof(char*));"
	print "\tfor (i = 0; i < s->n_frag; ++i) {"
	print "\t\tint off = s->seg_off[i];"
	print "\t\tfor (j = 0; j < s->n_seg[i]; ++j) { // as worker_for(), gmap.c:38-43"
	print "\t\t\tif (s->n_seg[i] == 2 && ((j == 0 && (pe_ori>>1&1)) || (j == 1 && (pe_ori&1)))) mg_revcomp_bseq(&s->seq[off + j]);"
	print "\t\t\tql[off + j] = s->seq[off + j].l_seq, sq[off + j] = s->seq[off + j].seq;"
	print "\t\t}"
	print "\t\tnm[i] = s->seq[off].name;"
	print "\t}"
	print "\tif (mg_map_batch_frag(s->p->gi, s->n_frag, s->n_seg, ql, sq, nm, s->gcs, s->p->opt) < 0) abort();"
	print "\tfree(ql); free(sq); free(nm);"
	print "}"
	print ""
}
/kt_for\(p->n_threads, worker_for, in, \(\(step_t\*\)in\)->n_frag\);/ {
	print "\t\tif (!(p->opt->flag & MG_M_INDEPEND_SEG)) mgb_worker_batch((step_t*)in);"
	print "\t\telse kt_for(p->n_threads, worker_for, in, ((step_t*)in)->n_frag);"
	next
}
{ print }
