/* mex.h -- MINIMAL STUB of the MATLAB MEX C API, for a SYNTAX CHECK ONLY.
 * MATLAB, Octave and the real mex.h are absent from this image (SURVEY.md 8(c) probe), so
 * admm_mex.cpp cannot be built into a MEX file or executed here.  This header declares exactly the
 * subset of the documented MEX API that the gateway uses, so that `g++ -fsyntax-only` can check the
 * gateway against include/admm_b200.h.  It implements nothing. */
#ifndef ADMMB_STUB_MEX_H
#define ADMMB_STUB_MEX_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef int mxClassIDStub;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxUNKNOWN_CLASS = 0, mxDOUBLE_CLASS = 6, mxINT32_CLASS = 12, mxINT64_CLASS = 14 } mxClassID;
bool mxIsStruct(const mxArray *);
bool mxIsDouble(const mxArray *);
bool mxIsInt32(const mxArray *);
bool mxIsChar(const mxArray *);
bool mxIsComplex(const mxArray *);
bool mxIsEmpty(const mxArray *);
mxArray *mxGetField(const mxArray *, size_t index, const char *name);
double *mxGetDoubles(const mxArray *);
int32_t *mxGetInt32s(const mxArray *);
int64_t *mxGetInt64s(const mxArray *);
double mxGetScalar(const mxArray *);
size_t mxGetNumberOfElements(const mxArray *);
mwSize mxGetNumberOfDimensions(const mxArray *);
const mwSize *mxGetDimensions(const mxArray *);
int mxGetString(const mxArray *, char *buf, mwSize buflen);
mxArray *mxCreateUninitNumericArray(size_t ndim, size_t *dims, mxClassID classid, mxComplexity flag);
mxArray *mxCreateNumericMatrix(size_t m, size_t n, mxClassID classid, mxComplexity flag);
mxArray *mxCreateStructMatrix(size_t m, size_t n, int nfields, const char **names);
void mxSetField(mxArray *, size_t index, const char *name, mxArray *value);
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...);
int mexAtExit(void (*fn)(void));
void mexLock(void);
void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);
#ifdef __cplusplus
}
#endif
#endif
