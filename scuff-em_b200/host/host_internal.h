// internal helpers of the host-side mirror (not installed)
#pragma once
#include <cmath>
#include <mutex>
#include <cstdio>
#include <cstring>
#include <strings.h>
#include "../../include/libscuff_b200.h"

#define MAXSTR 10000   /* long TRANS lines (many surfaces) must fit: the reference tolerates them */
#define MAXTOK 1000

namespace scuff {

void AddLiveDeviceObject(int delta);   // bookkeeping for SetCudaDevice (util.cc)
int Tokenize(char *s, char **Tokens, int MaxTokens, const char *Separators = " \t\r\n");
char *vstrdup(const char *fmt, ...);

// complex expression evaluator for MATERIAL sections (stands in for the reference's cmatheval)
struct CExpr;
CExpr *CExprCreate(const char *text, char **ErrMsg);
void CExprSetVar(CExpr *e, const char *name, cdouble value);
cdouble CExprEval(CExpr *e, cdouble w);
std::recursive_mutex &CExprMutex();        // held by CExprEval; hold it across CExprSetVar ... CExprEval sequences

inline double VecDistance(const double *a, const double *b)
{ double d = 0.0; for (int n = 0; n < 3; n++) d += (a[n] - b[n]) * (a[n] - b[n]); return sqrt(d); }

// rigid transformation X -> M X + DX (libs/libscuff/GTransformation.cc)
struct GTrans
{ double DX[3], M[3][3];
  GTrans() { Reset(); }
  void Reset() { memset(DX, 0, sizeof DX); memset(M, 0, sizeof M); M[0][0] = M[1][1] = M[2][2] = 1.0; }
  void Displace(const double dx[3]) { for (int i = 0; i < 3; i++) DX[i] += 1.0 * dx[i]; }
  void Rotate(const double ZHat[3], double ThetaDegrees);
  void Apply(double *X, int NX = 1) const;
  void UnApply(double *X, int NX = 1) const;
  void Transform(const GTrans *G);   // this = G o this
  bool Parse(char **Tokens, int NumTokens, char **ErrMsg, int *TokensConsumed);
};

}  // namespace scuff
