/*
 * ORACLE (test infrastructure, NOT product code).
 *
 * First pass of the classical Ruge-Stueben C/F colouring -- what HYPRE's coarsen type 10 (HMIS), the value the reference
 * requests (subdomain.tpp:1853) and HYPRE's default for the second setup (subdomain.tpp:3480-3489), does on a matrix that
 * lives on ONE process: HMIS runs the first Ruge-Stueben pass on the points without off-process connections and PMIS on the
 * rest, and on MPI_COMM_SELF there is no rest.   [Ruge, Stueben 1987; De Sterck, Yang, Heys 2006, section 3]
 * HYPRE itself is absent (parity unpinned, see oracle/amg.py); the tie-break between points of equal measure is made
 * explicit here: first in, first out, a point that changes measure re-entering at the tail of its new queue.
 *
 * Written independently of the product's linked-list version (csrc/host/amg.hpp): queues are arrays with lazy deletion --
 * every entry carries the stamp its point had when it was queued and is skipped if the point has moved since.
 */
#include <stdlib.h>
#include <string.h>

typedef struct { int *pt; int *stamp; int head, size, cap; } queue_t;

static void q_push(queue_t *q, int p, int stamp)
{
    if (q->size == q->cap)
    {
        q->cap = q->cap ? 2 * q->cap : 16;
        q->pt = (int *)realloc(q->pt, sizeof(int) * (size_t)q->cap);
        q->stamp = (int *)realloc(q->stamp, sizeof(int) * (size_t)q->cap);
    }
    q->pt[q->size] = p;
    q->stamp[q->size] = stamp;
    q->size++;
}

/* S: row i lists the points i strongly depends on.  cf out: +1 C, -1 F. */
void oracle_rs_first_pass(int n, const int *Sptr, const int *Scol, signed char *cf)
{
    int i, j, k, jj, kk;
    if (n <= 0) return;
    int *Tptr = (int *)calloc((size_t)n + 1, sizeof(int));
    int *Tcol = (int *)malloc(sizeof(int) * (size_t)(Sptr[n] > 0 ? Sptr[n] : 1));
    int *fill = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int *lambda = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int *stamp = (int *)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
    int maxl = 0, nq, top;
    queue_t *Q;
    for (i = 0; i < Sptr[n]; i++) Tptr[Scol[i] + 1]++;
    for (i = 0; i < n; i++) Tptr[i + 1] += Tptr[i];
    for (i = 0; i < n; i++) fill[i] = Tptr[i];
    for (i = 0; i < n; i++)
        for (j = Sptr[i]; j < Sptr[i + 1]; j++) Tcol[fill[Scol[j]]++] = i;   /* S^T: who depends on me */
    for (i = 0; i < n; i++) { lambda[i] = Tptr[i + 1] - Tptr[i]; if (lambda[i] > maxl) maxl = lambda[i]; }
    nq = 2 * maxl + 2;
    Q = (queue_t *)calloc((size_t)nq, sizeof(queue_t));
    memset(cf, 0, (size_t)n);
    top = 0;
    for (i = 0; i < n; i++)
    {
        if (lambda[i] == 0) { cf[i] = -1; continue; }
        q_push(&Q[lambda[i]], i, 0);
        if (lambda[i] > top) top = lambda[i];
    }
    for (;;)
    {
        queue_t *q;
        int pick = -1;
        while (top > 0)
        {
            q = &Q[top];
            while (q->head < q->size)
            {
                int p = q->pt[q->head];
                if (cf[p] == 0 && stamp[p] == q->stamp[q->head] && lambda[p] == top) { pick = p; break; }
                q->head++;                                               /* stale entry */
            }
            if (pick >= 0) break;
            top--;
        }
        if (pick < 0) break;
        Q[top].head++;
        cf[pick] = 1;
        for (jj = Tptr[pick]; jj < Tptr[pick + 1]; jj++)
        {
            j = Tcol[jj];
            if (cf[j] != 0) continue;
            cf[j] = -1;                                                  /* depends strongly on the new C point */
            for (kk = Sptr[j]; kk < Sptr[j + 1]; kk++)
            {
                k = Scol[kk];
                if (cf[k] != 0) continue;
                lambda[k]++; stamp[k]++;
                q_push(&Q[lambda[k]], k, stamp[k]);
                if (lambda[k] > top) top = lambda[k];
            }
        }
        for (jj = Sptr[pick]; jj < Sptr[pick + 1]; jj++)
        {
            j = Scol[jj];
            if (cf[j] != 0) continue;
            lambda[j]--; stamp[j]++;
            if (lambda[j] <= 0) cf[j] = -1;
            else q_push(&Q[lambda[j]], j, stamp[j]);
        }
    }
    for (i = 0; i < nq; i++) { free(Q[i].pt); free(Q[i].stamp); }
    free(Q); free(Tptr); free(Tcol); free(fill); free(lambda); free(stamp);
}
