/* hostio.h — pageable-host transfers of the reference-facing entry points (see hostio.c). */
#ifndef QRDM_HOSTIO_H_
#define QRDM_HOSTIO_H_
#ifdef __cplusplus
extern "C" {
#endif
typedef struct qrdm_hostio qrdm_hostio;
int qrdm_hostio_create(qrdm_hostio **io, int device);
void qrdm_hostio_destroy(qrdm_hostio *io);
/* blocking, multi-threaded bounce copies of an m x n column-major matrix (columns [c0, n) for download) */
int qrdm_hostio_upload(qrdm_hostio *io, double *d, int ldd, const double *h, int ldh, int m, int n);
int qrdm_hostio_download(qrdm_hostio *io, double *h, int ldh, const double *d, int ldd, int m, int c0, int n);
/* streamed write-back: begin returns 0 (active), 1 (not worthwhile for this shape: inactive), < 0 error */
int qrdm_hostio_wb_begin(qrdm_hostio *io, double *h, int ldh, const double *d, int ldd, int m, void *copy_stream);
int qrdm_hostio_wb_push(qrdm_hostio *io, int c0, int c1); /* returns columns taken (prefix of the range) */
int qrdm_hostio_wb_end(qrdm_hostio *io);                  /* waits until every taken column is in the host buffer */
#ifdef __cplusplus
}
#endif
#endif
