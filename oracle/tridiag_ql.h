/* TEST INFRASTRUCTURE ONLY -- see tridiag_ql.c */
#ifndef CUPPEN_ORACLE_TRIDIAG_QL_H
#define CUPPEN_ORACLE_TRIDIAG_QL_H
#ifdef __cplusplus
extern "C" {
#endif
int cuppen_oracle_tql2(int n, double *d, const double *e, double *z, int ldz, int row_major);
#ifdef __cplusplus
}
#endif
#endif
