/*
 * fpo_mp.c -- multi-worker driver of the oracle for the CPU baseline
 * (test infrastructure; used only by bench.py's cpu_baseline /
 * --impl reference legs and tests).
 *
 * Mirrors the FLEXPART_MPI execution model (README_PARALLEL.md:60-73,
 * src/mpi_mod.f90:323, src/timemanager_mpi.f90:468): P workers, each with
 * its own copy of the module state and its own share of the particles,
 * sharing the read-only met arrays; grids are summed by the caller at output
 * time (mpif_tm_reduce_grid).  Workers are POSIX threads instead of MPI ranks.
 */
#include <pthread.h>
#include <stdlib.h>
#include <time.h>

#include "fpo.h"

typedef struct {
  fpo_state *S;
  int itime, ldeltat;
  float w;
} job_t;

static void *worker(void *arg) {
  job_t *j = (job_t *)arg;
  if (j->w > 0.f) fpo_conccalc(j->S, j->itime, j->w);
  fpo_step(j->S, j->itime, j->ldeltat, NULL);
  return NULL;
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* run fpo_conccalc (when weight > 0) + fpo_step on every state, one thread
 * per state; returns the wall time in seconds. */
double fpo_mp_step(fpo_state **states, int nstates, int itime, int ldeltat,
                   float conc_weight) {
  pthread_t *th = (pthread_t *)calloc((size_t)nstates, sizeof(pthread_t));
  job_t *jobs = (job_t *)calloc((size_t)nstates, sizeof(job_t));
  double t0 = now_s();
  for (int s = 0; s < nstates; s++) {
    jobs[s].S = states[s];
    jobs[s].itime = itime;
    jobs[s].ldeltat = ldeltat;
    jobs[s].w = conc_weight;
    pthread_create(&th[s], NULL, worker, &jobs[s]);
  }
  for (int s = 0; s < nstates; s++) pthread_join(th[s], NULL);
  double t1 = now_s();
  free(th);
  free(jobs);
  return t1 - t0;
}
