/*
 * emub_snapshot.h -- reader for the reference's MODEL_SNAPSHOT_FILE (ASCII), plain C.
 * Grammar: dump_multi_modelstruct (src/multi_modelstruct.c:346-401) followed by one dump_modelstruct_2 block
 * per PCA component (src/modelstruct.c:375-409); read exactly as load_multi_modelstruct (:406-472) and
 * load_modelstruct_2 (modelstruct.c:419-467) do.  The format itself is untouched.
 */
#ifndef EMUB_SNAPSHOT_H
#define EMUB_SNAPSHOT_H
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	int nthetas, nparams, nmodel_points, nemulate_points, regression_order, nregression_fns;
	int fixed_nugget_mode, cov_fn_index, use_data_scales;
	double fixed_nugget;
	double *grad_ranges;     /* nthetas x 2 */
	double *xmodel;          /* n x d */
	double *training_vector; /* n */
	double *thetas;          /* nthetas */
	double *sample_scales;   /* d */
} emub_snapshot_component;

typedef struct {
	int nt, nr, nparams, nmodel_points, cov_fn_index, regression_order;
	double *xmodel;          /* n x d */
	double *training_matrix; /* n x nt */
	double *training_mean;   /* nt, recomputed as the loader does (multi_modelstruct.c:464-469) */
	double *pca_evals_r;     /* nr */
	double *pca_evecs_r;     /* nt x nr */
	double *pca_zmatrix;     /* n x nr */
	emub_snapshot_component *components; /* nr */
} emub_snapshot;

/* returns NULL (and a message in err, if given) on a malformed file */
emub_snapshot *emub_snapshot_load(FILE *f, char *err, int errlen);
emub_snapshot *emub_snapshot_load_path(const char *path, char *err, int errlen);
/* writer: the bytes dump_multi_modelstruct (multi_modelstruct.c:346-401) + dump_modelstruct_2 (modelstruct.c:375-409)
 * produce for the same contents; 0 on success */
int emub_snapshot_save(const emub_snapshot *s, FILE *f);
int emub_snapshot_save_path(const emub_snapshot *s, const char *path);
/* nr scalar GPs on one design as a snapshot with an identity back-projection (nt = nr, U = I, lambda = 1):
 * Z n x nr, thetas nr x nthetas (full vectors, amplitude first) */
emub_snapshot *emub_snapshot_from_arrays(const double *X, int n, int d, const double *Z, int nr, const double *thetas, int nthetas,
                                         int cov_fn_index, int regression_order);
void emub_snapshot_free(emub_snapshot *s);

#ifdef __cplusplus
}
#endif
#endif
