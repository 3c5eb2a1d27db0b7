// Material properties eps(w), mu(w) evaluated on the host and passed to the device by value.
// Mirrors libs/libMatProp/libMatProp.cc:160-300 (name conventions PEC / VACUUM / CONST_EPS_x[_MU_y]
// [_SIGMA_z]), :501-553 (GetEpsMu) and UserDefinedMaterials.cc:140-392 (MATERIAL...ENDMATERIAL
// sections with constants and Eps(w)/Mu(w) expressions).  The reference evaluates expressions with
// its bundled cmatheval; here a small recursive-descent complex evaluator stands in for it.
// Tabulated (FILE_...) materials are not supported.
#include <mutex>
#include <cctype>
#include <cstdlib>
#include <map>
#include "host_internal.h"

namespace scuff {

double MatProp::FreqUnit = 2.99792458e14;   // c / (1 micron), libMatProp.cc:44

// ---------------------------------------------------------------------------------------
// expression evaluator: + - * / ^, unary -, parentheses, numbers, i/I, variables, a few functions
// ---------------------------------------------------------------------------------------
struct CExpr
{ std::string text;
  std::map<std::string, cdouble> vars;
  // parser state
  const char *p; cdouble w; bool ok; std::string err;
  void skip() { while (*p && isspace((unsigned char)*p)) p++; }
  cdouble parseExpr()
  { cdouble v = parseTerm();
    for (;;)
     { skip();
       if (*p == '+') { p++; v = v + parseTerm(); }
       else if (*p == '-') { p++; v = v - parseTerm(); }
       else return v;
     }
  }
  cdouble parseTerm()
  { cdouble v = parseUnary();
    for (;;)
     { skip();
       if (*p == '*') { p++; v = v * parseUnary(); }
       else if (*p == '/') { p++; v = v / parseUnary(); }
       else return v;
     }
  }
  cdouble parseUnary()
  { skip();
    if (*p == '-') { p++; return -parseUnary(); }
    if (*p == '+') { p++; return parseUnary(); }
    return parsePower();
  }
  cdouble parsePower()
  { cdouble b = parsePrimary();
    skip();
    if (*p == '^') { p++; cdouble e = parseUnary(); return std::pow(b, e); }
    return b;
  }
  cdouble parsePrimary()
  { skip();
    if (*p == '(')
     { p++; cdouble v = parseExpr(); skip();
       if (*p == ')') p++; else { ok = false; err = "missing )"; }
       return v; }
    if (isdigit((unsigned char)*p) || *p == '.')
     { char *end; double d = strtod(p, &end); p = end; return cdouble(d, 0.0); }
    if (isalpha((unsigned char)*p) || *p == '_')
     { std::string id;
       while (isalnum((unsigned char)*p) || *p == '_') id += *p++;
       skip();
       if (*p == '(')
        { p++; cdouble a = parseExpr(); skip();
          if (*p == ')') p++; else { ok = false; err = "missing )"; }
          std::string f = id; for (auto &c : f) c = tolower(c);
          if (f == "sqrt") return std::sqrt(a);
          if (f == "exp") return std::exp(a);
          if (f == "log") return std::log(a);
          if (f == "sin") return std::sin(a);
          if (f == "cos") return std::cos(a);
          if (f == "tan") return std::tan(a);
          if (f == "sinh") return std::sinh(a);
          if (f == "cosh") return std::cosh(a);
          if (f == "tanh") return std::tanh(a);
          if (f == "abs") return cdouble(std::abs(a), 0.0);
          if (f == "real") return cdouble(real(a), 0.0);
          if (f == "imag") return cdouble(imag(a), 0.0);
          if (f == "conj") return std::conj(a);
          ok = false; err = "unknown function " + id; return 0.0;
        }
       if (id == "w") return w;
       auto it = vars.find(id);
       if (it != vars.end()) return it->second;
       if (id == "i" || id == "I") return cdouble(0.0, 1.0);
       if (id == "pi" || id == "Pi" || id == "PI") return cdouble(M_PI, 0.0);
       ok = false; err = "unknown symbol " + id; return 0.0;
     }
    ok = false; err = std::string("unexpected character '") + *p + "'";
    if (*p) p++;
    return 0.0;
  }
  cdouble eval(cdouble ww)
  { p = text.c_str(); w = ww; ok = true;
    cdouble v = parseExpr(); skip();
    if (*p) { ok = false; err = "junk at end of expression"; }
    return v;
  }
};

CExpr *CExprCreate(const char *text, char **ErrMsg)
{ CExpr *e = new CExpr; e->text = text;
  if (ErrMsg) *ErrMsg = 0;
  return e;
}
void CExprSetVar(CExpr *e, const char *name, cdouble value) { e->vars[name] = value; }
// Parsed expressions are shared between the geometries of a process (a MATERIAL section is kept once in the material data base, as in
// the reference) and evaluation mutates the expression object (parser position, error state): worker threads of a SweepDriver evaluating
// Eps(w) for different frequencies at the same time raced here.  One lock for all expression objects; callers that set variables and then
// evaluate (SURFACE_IMPEDANCE) hold it across the sequence.
std::recursive_mutex &CExprMutex() { static std::recursive_mutex m; return m; }

cdouble CExprEval(CExpr *e, cdouble w)
{ std::lock_guard<std::recursive_mutex> lock(CExprMutex());
  cdouble v = e->eval(w);
  if (!e->ok) ErrExit("invalid material expression '%s': %s", e->text.c_str(), e->err.c_str());
  return v;
}

// ---------------------------------------------------------------------------------------
// database of user-defined materials (MATERIAL sections in .scuffgeo files)
// ---------------------------------------------------------------------------------------
struct UserMaterial { std::string name; CExpr *eps, *mu; };
static std::vector<UserMaterial> g_userMaterials;

char *AddMaterialToMatPropDataBase(FILE *f, const char *FileName, const char *MaterialName, int *LineNum)
{ UserMaterial UM; UM.name = MaterialName; UM.eps = 0; UM.mu = 0;
  std::vector<std::pair<std::string, double> > constants;
  char Line[MAXSTR];
  while (fgets(Line, MAXSTR, f))
   { (*LineNum)++;
     char *p = Line; while (*p && isspace((unsigned char)*p)) p++;
     if (*p == 0 || *p == '#') continue;
     if (!strncasecmp(p, "ENDMATERIAL", 11))
      { // replace an earlier definition of the same name, as a later MATERIAL section would shadow it
        for (auto &m : g_userMaterials) if (!strcasecmp(m.name.c_str(), MaterialName)) { m = UM; return 0; }
        g_userMaterials.push_back(UM); return 0; }
     bool isEps = !strncasecmp(p, "Eps(w)", 6), isMu = !strncasecmp(p, "Mu(w)", 5);
     char *pp = strchr(p, '=');
     if (!pp) return vstrdup("%s:%i: %s", FileName, *LineNum, (isEps || isMu) ? "syntax error" : "unrecognized input");
     if (isEps || isMu)
      { char *ppp = strrchr(pp, ';');
        if (ppp) *ppp = 0; else { ppp = strrchr(pp, '\n'); if (ppp) *ppp = 0; }
        CExpr *e = CExprCreate(pp + 1, 0);
        for (auto &c : constants) CExprSetVar(e, c.first.c_str(), cdouble(c.second, 0.0));
        e->eval(cdouble(1.0e14, 0.0));
        if (!e->ok) return vstrdup("%s:%i: invalid expression", FileName, *LineNum);
        if (isEps) UM.eps = e; else UM.mu = e;
      }
     else
      { double CV; int n;
        if (sscanf(pp + 1, "%le%n", &CV, &n) < 1) return vstrdup("%s:%i: invalid constant value", FileName, *LineNum);
        char *ppp = pp + 1 + n;
        while (*ppp != 0 && *ppp != '#' && (isspace((unsigned char)*ppp) || *ppp == ';')) ppp++;
        if (*ppp != 0 && *ppp != '#') return vstrdup("%s:%i: junk at end of line", FileName, *LineNum);
        while (pp > p && isspace((unsigned char)*(pp - 1))) pp--;
        *pp = 0;
        constants.push_back(std::make_pair(std::string(p), CV));
      }
   }
  return vstrdup("%s: unexpected end of file", FileName);
}

// ---------------------------------------------------------------------------------------
enum { MP_PEC = 0, MP_VACUUM = 1, MP_CONSTANT = 2, MP_INTERP = 3, MP_PARSED = 4 };

MatProp::MatProp(const char *MaterialName)
{ ErrMsg = 0; Zeroed = 0; Eps = 1.0; Mu = 1.0; EpsExpr = MuExpr = 0;
  const char *p = MaterialName;
  if (p == 0 || *p == 0 || !strcasecmp(p, "VACUUM")) { Type = MP_VACUUM; Name = strdup(p && *p ? p : "VACUUM"); return; }
  Name = strdup(p);
  if (!strcasecmp(p, "PEC")) { Type = MP_PEC; return; }
  if (!strncasecmp(p, "CONST_", 6))
   { char ConstString[100]; char *Tokens[6]; double ER, EI; char c1, c2; double Sigma = 0.0;
     strncpy(ConstString, p + 6, 99); ConstString[99] = 0;
     int NumTokens = Tokenize(ConstString, Tokens, 6, "_");
     for (int nt = 0; nt < NumTokens; nt++)
      { bool isMu = !strcasecmp(Tokens[nt], "MU"), isEps = !strcasecmp(Tokens[nt], "EPS");
        if (isMu || isEps)
         { if (++nt == NumTokens) { ErrMsg = strdup(isMu ? "no value specified for MU" : "no value specified for EPS"); return; }
           int nConv = sscanf(Tokens[nt], "%le%c%le%c", &ER, &c1, &EI, &c2);
           cdouble val;
           if (nConv == 1) val = cdouble(ER, 0.0);
           else if (nConv == 4 && c1 == '+' && tolower(c2) == 'i') val = cdouble(ER, EI);
           else if (nConv == 4 && c1 == '-' && tolower(c2) == 'i') val = cdouble(ER, -1.0 * EI);
           else { ErrMsg = strdup(isMu ? "bad constant value specified for MU" : "bad constant value specified for EPS"); return; }
           if (isMu) Mu = val; else Eps = val;
         }
        else if (!strcasecmp(Tokens[nt], "SIGMA"))
         { if (++nt == NumTokens) { ErrMsg = strdup("no value specified for SIGMA"); return; }
           if (sscanf(Tokens[nt], "%le", &Sigma) != 1) { ErrMsg = strdup("bad constant value specified for Sigma"); return; }
         }
        else { ErrMsg = strdup("invalid token"); return; }
      }
     if (Sigma == 0.0) Type = MP_CONSTANT;
     else
      { Type = MP_PARSED;
        char Expr[100];
        // same text (and hence the same %e-rounded constants) as libMatProp.cc:277
        snprintf(Expr, 100, "%e + %e*i*2.99792458e14/w", real(Eps), Sigma);
        EpsExpr = CExprCreate(Expr, 0);
        if (Mu != 1.0) { snprintf(Expr, 100, "%e+%e*i", real(Mu), imag(Mu)); MuExpr = CExprCreate(Expr, 0); }
      }
     return;
   }
  for (auto &m : g_userMaterials)
   if (!strcasecmp(m.name.c_str(), p)) { Type = MP_PARSED; EpsExpr = m.eps; MuExpr = m.mu; return; }
  ErrMsg = vstrdup("could not find material %s (tabulated / database materials are not supported)", p);
  Type = MP_VACUUM;
}

void MatProp::GetEpsMu(cdouble Omega, cdouble *pEps, cdouble *pMu)
{ cdouble EpsRV, MuRV;
  if (Zeroed) { EpsRV = 0.0; MuRV = 0.0; }
  else if (Type == MP_VACUUM) { EpsRV = 1.0; MuRV = 1.0; }
  else if (Type == MP_PEC) { EpsRV = -1.0e9; MuRV = 1.0; }
  else if (Type == MP_CONSTANT) { EpsRV = Eps; MuRV = Mu; }
  else
   { cdouble w = Omega * FreqUnit;
     EpsRV = EpsExpr ? CExprEval((CExpr *)EpsExpr, w) : cdouble(1.0, 0.0);
     MuRV = MuExpr ? CExprEval((CExpr *)MuExpr, w) : cdouble(1.0, 0.0);
   }
  if (pEps) *pEps = EpsRV;
  if (pMu) *pMu = MuRV;
}

}  // namespace scuff
