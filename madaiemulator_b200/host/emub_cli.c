/*
 * emub_cli.c -- `emub_interactive_emulator interactive_mode MODEL_SNAPSHOT_FILE [--quiet] [--pca_output]`:
 * the reference CLI's interactive_mode (src/interactive_emulator.c:369-450, usage :160-206) served by the B200
 * engine with block I/O.  Reads points from stdin, writes the reference's text protocol to stdout.
 * (estimate_thetas / print_thetas stay with the reference CLI, bound through integration/libemu_glue.c.)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "emub_interactive.h"

int main(int argc, char **argv)
{
	int quiet = 0, pca = 0, binary = 0, block = 0;
	int devices[64] = {0}, ndev = 1;
	const char *snap = NULL;
	if (argc < 3 || strcmp(argv[1], "interactive_mode") != 0) {
		fprintf(stderr, "usage: %s interactive_mode MODEL_SNAPSHOT_FILE [--quiet|-q] [--pca_output] [--binary] [--devices 0,1,..] [--block N]\n", argv[0]);
		return 2;
	}
	snap = argv[2];
	for (int i = 3; i < argc; i++) {
		if (!strcmp(argv[i], "--quiet") || !strcmp(argv[i], "-q")) quiet = 1;
		else if (!strcmp(argv[i], "--pca_output")) { pca = 1; quiet = 1; } /* the reference's missing break, interactive_emulator.c:589-601 */
		else if (!strcmp(argv[i], "--binary")) binary = 1;
		else if ((!strcmp(argv[i], "--device") || !strcmp(argv[i], "--devices")) && i + 1 < argc) {
			/* comma separated list: one replica per GPU, query blocks are shared out between them */
			char *tok = strtok(argv[++i], ",");
			for (ndev = 0; tok && ndev < 64; tok = strtok(NULL, ",")) devices[ndev++] = atoi(tok);
			if (ndev == 0) { devices[0] = 0; ndev = 1; }
		}
		else if (!strcmp(argv[i], "--block") && i + 1 < argc) block = atoi(argv[++i]);
	}
	char err[256] = "";
	emub_snapshot *s = emub_snapshot_load_path(snap, err, sizeof(err));
	if (!s) { fprintf(stderr, "%s: %s\n", snap, err); return 1; }
	emub_multi_emulator *me = NULL;
	if (emub_multi_emulator_from_snapshot_devices(devices, ndev, s, &me) != EMUB_OK) { fprintf(stderr, "%s\n", emub_last_error()); return 1; }
	long long npts = 0;
	int rc = emub_interactive_stream(me, stdin, stdout, quiet, pca, binary, block, &npts);
	if (rc != EMUB_OK) fprintf(stderr, "%s\n", emub_last_error());
	emub_multi_emulator_destroy(me);
	emub_snapshot_free(s);
	return rc == EMUB_OK ? 0 : 1;
}
