# Applied to the reference's gmap.c (read in place from /root/reference, never copied into the repo): produces
# oracle/_ref/gmap_b200.c, the ONLY host file that changes when minigraph is linked against libmgb200.so.
# (1) a 12-line helper that hands one whole mini-batch to the GPU dispatcher, (2) the kt_for call at gmap.c:99.
/^static void \*worker_pipeline\(void \*shared, int step, void \*in\)/ {
	print "extern int mg_map_batch(const mg_idx_t *gi, int n_reads, const int *qlens, const char *const *seqs, const char *const *names, mg_gchains_t **gcs, const mg_mapopt_t *opt);"
	print "static void mgb_worker_batch(step_t *s) // replaces kt_for(worker_for) when every fragment is one read"
	print "{"
	print "\tint i, n = s->n_seq, *ql = (int*)malloc(n * sizeof(int));"
	print "\tconst char **sq = (const char**)malloc(n * sizeof(char*)), **nm = (const char**)malloc(n * sizeof(char*));"
	print "\tfor (i = 0; i < n; ++i) ql[i] = s->seq[i].l_seq, sq[i] = s->seq[i].seq, nm[i] = s->seq[i].name;"
	print "\tif (mg_map_batch(s->p->gi, n, ql, sq, nm, s->gcs, s->p->opt) < 0) abort();"
	print "\tfree(ql); free(sq); free(nm);"
	print "}"
	print ""
}
/kt_for\(p->n_threads, worker_for, in, \(\(step_t\*\)in\)->n_frag\);/ {
	print "\t\tif (((step_t*)in)->n_frag == ((step_t*)in)->n_seq && !(p->opt->flag & MG_M_INDEPEND_SEG)) mgb_worker_batch((step_t*)in);"
	print "\t\telse kt_for(p->n_threads, worker_for, in, ((step_t*)in)->n_frag);"
	next
}
{ print }
