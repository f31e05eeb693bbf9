/*
 * smart_b200_io.h -- host-side bulk text I/O of the sample database (libsmart_b200.so).
 *
 * What it replaces.  The reference writes its Monte-Carlo database one line per sample from inside
 * spotpy's loop -- `'%.6e'` of the float32 rounding of every value, comma separated
 * (smartpy/montecarlo/montecarlo.py:211-231, header :125-127) -- and reads it back row by row
 * through csv.DictReader into float32 arrays (:233-262).  At 1e7 samples that text is the step
 * right after the hot path: numpy.savetxt needs ~9 us per row of 18
 * values.  The two functions below format and parse whole blocks of rows on all host cores and
 * produce / accept exactly the same bytes.  Plain C ABI, host pointers only, no CUDA involved.
 */
#ifndef SMART_B200_IO_H
#define SMART_B200_IO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Bytes that always hold n_rows x n_cols formatted values ("-d.dddddde+dd" is 13 characters, plus
 * a separator each). */
int64_t smart_csv_bound(int64_t n_rows, int32_t n_cols);

/*
 * table[n_rows][ld] (float32, the first n_cols columns of each row) -> text in out: every value as
 * '%.6e' of the value (Python / numpy.savetxt semantics: 'nan', 'inf', '-inf' for the non-finite
 * ones), ',' between values, '\n' after each row.  n_threads <= 0: all host cores.
 * Returns the number of bytes written, or a negative SMART_ERR_* code (bad argument, out_cap below
 * smart_csv_bound()).
 */
int64_t smart_csv_format_f32(const float *table, int64_t n_rows, int32_t n_cols, int64_t ld, char *out,
                             int64_t out_cap, int32_t n_threads);

/*
 * The inverse, for the columns a caller wants: text[n_bytes] holds lines of n_cols_in_file
 * comma-separated numbers (no header line); out[row][k] = float32 of the binary64 value of column
 * wanted[k] (text -> binary64 -> float32, the conversion numpy makes of the reference's list of
 * strings).  Returns the number of rows parsed (<= max_rows), or a negative code: SMART_ERR_BAD_ARG
 * for a bad argument or a line that does not hold n_cols_in_file numbers.
 */
int64_t smart_csv_parse_f32(const char *text, int64_t n_bytes, int32_t n_cols_in_file, const int32_t *wanted,
                            int32_t n_wanted, float *out, int64_t max_rows, int32_t n_threads);

#ifdef __cplusplus
}
#endif
#endif
