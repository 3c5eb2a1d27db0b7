// Geometry front-end with the reference's API surface: .scuffgeo / GMSH .msh parsing, RWG edge
// list construction (this DEFINES the row/column ordering of the BEM matrix), region / nesting
// bookkeeping, Mate[] detection, transformations, and flattening to the device (C ABI).
//
// Reference semantics restated here (not copied): RWGGeometry::RWGGeometry RWGGeometry.cc:353-806,
// AddRegion :176-215, ProcessLATTICESection :126-171; RWGSurface ctor/InitRWGSurface RWGSurface.cc:
// 263-529, InitRWGPanel :55-97, Transform/UnTransform :588-667; InitEdgeList InitEdgeList.cc:47-403;
// ParseGMSHFile ParseMeshFiles.cc:69-246; GTransformation.cc:96-339.
#include <cstdarg>
#include <cstdlib>
#include <algorithm>
#include <atomic>
#include "host_internal.h"

namespace scuff {

bool RWGGeometry::UseHRWGFunctions = false;
bool RWGGeometry::UseHighKTaylorDuffy = true;
bool RWGGeometry::UseTaylorDuffyV2P0 = true;
bool RWGGeometry::DisableCache = false;

// ---------------------------------------------------------------------------------------
// GTrans
// ---------------------------------------------------------------------------------------
static void RotateVec(const double U[3], double c, double s, double V[3])
{ double d = 0.0; for (int i = 0; i < 3; i++) d += U[i] * V[i];
  double Vpar[3], Vperp[3], Vcross[3];
  for (int i = 0; i < 3; i++) Vpar[i] = U[i] * d;
  for (int i = 0; i < 3; i++) Vperp[i] = V[i] - Vpar[i];
  Vcross[0] = U[1] * V[2] - U[2] * V[1]; Vcross[1] = U[2] * V[0] - U[0] * V[2]; Vcross[2] = U[0] * V[1] - U[1] * V[0];
  for (int i = 0; i < 3; i++) V[i] = Vpar[i] + (Vperp[i] * c + Vcross[i] * s);
}
void GTrans::Rotate(const double ZHat[3], double ThetaDegrees)
{ double U[3] = { ZHat[0], ZHat[1], ZHat[2] };
  double n = sqrt(U[0] * U[0] + U[1] * U[1] + U[2] * U[2]); double in = 1.0 / n;
  for (int i = 0; i < 3; i++) U[i] *= in;
  ThetaDegrees *= (3.14159265358979323846 / 180.0);
  double c = cos(ThetaDegrees), s = sin(ThetaDegrees);
  RotateVec(U, c, s, DX);
  for (int col = 0; col < 3; col++)
   { double t[3] = { M[0][col], M[1][col], M[2][col] }; RotateVec(U, c, s, t);
     M[0][col] = t[0]; M[1][col] = t[1]; M[2][col] = t[2]; }
}
void GTrans::Apply(double *X, int NX) const
{ for (int nx = 0; nx < NX; nx++)
   { double x0 = X[3 * nx], x1 = X[3 * nx + 1], x2 = X[3 * nx + 2];
     X[3 * nx + 0] = M[0][0] * x0 + M[0][1] * x1 + M[0][2] * x2 + DX[0];
     X[3 * nx + 1] = M[1][0] * x0 + M[1][1] * x1 + M[1][2] * x2 + DX[1];
     X[3 * nx + 2] = M[2][0] * x0 + M[2][1] * x1 + M[2][2] * x2 + DX[2]; }
}
void GTrans::UnApply(double *X, int NX) const
{ for (int nx = 0; nx < NX; nx++)
   { double x0 = X[3 * nx] - DX[0], x1 = X[3 * nx + 1] - DX[1], x2 = X[3 * nx + 2] - DX[2];
     X[3 * nx + 0] = M[0][0] * x0 + M[1][0] * x1 + M[2][0] * x2;
     X[3 * nx + 1] = M[0][1] * x0 + M[1][1] * x1 + M[2][1] * x2;
     X[3 * nx + 2] = M[0][2] * x0 + M[1][2] * x1 + M[2][2] * x2; }
}
void GTrans::Transform(const GTrans *G)
{ G->Apply(DX);
  double newM[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
   { newM[i][j] = 0; for (int k = 0; k < 3; k++) newM[i][j] += G->M[i][k] * M[k][j]; }
  memcpy(M, newM, sizeof newM);
}
bool GTrans::Parse(char **Tokens, int NumTokens, char **ErrMsg, int *TokensConsumed)
{ double dx[3], Theta, ZHat[3];
  if (ErrMsg) *ErrMsg = 0;
  if (NumTokens == 0) { if (ErrMsg) *ErrMsg = strdup("no tranformation specified"); if (TokensConsumed) *TokensConsumed = 0; return false; }
  if (!strcasecmp(Tokens[0], "DISP") || !strcasecmp(Tokens[0], "DISPLACED"))
   { if (NumTokens < 4) { if (ErrMsg) *ErrMsg = vstrdup("too few values specified for %s ", Tokens[0]); return false; }
     if (1 != sscanf(Tokens[1], "%le", dx) || 1 != sscanf(Tokens[2], "%le", dx + 1) || 1 != sscanf(Tokens[3], "%le", dx + 2))
      { if (ErrMsg) *ErrMsg = vstrdup("bad value specified for %s", Tokens[0]); return false; }
     if (TokensConsumed) *TokensConsumed = 4;
     Displace(dx); return true;
   }
  if (!strcasecmp(Tokens[0], "ROT") || !strcasecmp(Tokens[0], "ROTATED"))
   { if (NumTokens < 6) { if (ErrMsg) *ErrMsg = vstrdup("too few values specified for %s ", Tokens[0]); return false; }
     if (strcasecmp(Tokens[2], "ABOUT")) { if (ErrMsg) *ErrMsg = vstrdup("invalid syntax for %s statement", Tokens[0]); return false; }
     if (1 != sscanf(Tokens[1], "%le", &Theta) || 1 != sscanf(Tokens[3], "%le", ZHat) || 1 != sscanf(Tokens[4], "%le", ZHat + 1) || 1 != sscanf(Tokens[5], "%le", ZHat + 2))
      { if (ErrMsg) *ErrMsg = vstrdup("bad value specified for %s", Tokens[0]); return false; }
     if (TokensConsumed) *TokensConsumed = 6;
     Rotate(ZHat, Theta); return true;
   }
  if (ErrMsg) *ErrMsg = vstrdup("unknown keyword %s", Tokens[0]);
  if (TokensConsumed) *TokensConsumed = 0;
  return false;
}

// ---------------------------------------------------------------------------------------
// mesh file parsing (GMSH v1 "$NOD" and v2 "$Nodes" ASCII formats)
// ---------------------------------------------------------------------------------------
static FILE *OpenWithPath(const std::vector<std::string> &dirs, const char *name)
{ FILE *f = fopen(name, "r"); if (f) return f;
  for (auto &d : dirs) { std::string p = d + "/" + name; f = fopen(p.c_str(), "r"); if (f) return f; }
  const char *env = getenv("SCUFF_MESH_PATH");
  if (env)
   { std::string s = env; size_t pos = 0;
     while (pos <= s.size())
      { size_t q = s.find(':', pos); if (q == std::string::npos) q = s.size();
        if (q > pos) { std::string p = s.substr(pos, q - pos) + "/" + name; f = fopen(p.c_str(), "r"); if (f) return f; }
        pos = q + 1; } }
  return 0;
}
static std::vector<std::string> g_meshDirs;

static char *ParseGMSHFile(FILE *MeshFile, const char *FileName, int MeshTag,
                           std::vector<double> &VC, std::vector<int> &PVI)
{ int LineNum = 0; char Line[100]; bool found = false; bool legacy = false;
  while (!found)
   { if (!fgets(Line, 100, MeshFile)) return vstrdup("%s: failed to find node start keyword", FileName);
     LineNum++;
     if (!strncmp(Line, "$NOD", 4)) { legacy = true; found = true; }
     else if (!strncmp(Line, "$Nodes", 6)) { legacy = false; found = true; }
   }
  int NumNodes;
  if (!fgets(Line, 100, MeshFile) || !(NumNodes = atoi(Line))) return vstrdup("%s:%i: invalid number of nodes", FileName, LineNum);
  LineNum++;
  VC.clear();
  std::vector<int> GMSH2HR(2 * (size_t)NumNodes + 1, -1);
  for (int nn = 0; nn < NumNodes; nn++)
   { if (!fgets(Line, 100, MeshFile)) return vstrdup("%s: too few nodes", FileName);
     LineNum++;
     int NodeIndex; double V[3];
     if (sscanf(Line, "%i %le %le %le", &NodeIndex, V, V + 1, V + 2) != 4) return vstrdup("%s:%i: invalid node specification", FileName, LineNum);
     if (NodeIndex > 2 * NumNodes || NodeIndex < 0) return vstrdup("%s:%i: internal error in ReadGMSHFile", FileName, LineNum);
     VC.push_back(V[0]); VC.push_back(V[1]); VC.push_back(V[2]);
     GMSH2HR[NodeIndex] = nn;
   }
  const char *px = getenv("SCUFF_PIXEL_SIZE");
  if (px) { double ps; if (sscanf(px, "%le", &ps) == 1) for (auto &v : VC) v = ps * round(v / ps); }
  // merge redundant vertices (distance < 1e-6): later duplicates are remapped to the first
  for (int i = 0; i < NumNodes; i++)
   for (int j = i + 1; j < NumNodes; j++)
    if (VecDistance(&VC[3 * i], &VC[3 * j]) < 1.0e-6)
     for (size_t g = 0; g < GMSH2HR.size(); g++) if (GMSH2HR[g] == j) GMSH2HR[g] = i;
  if (!fgets(Line, 100, MeshFile)) return vstrdup("%s: bad file format (nodes section not terminated)", FileName);
  LineNum++;
  if (strncmp(Line, legacy ? "$ENDNOD" : "$EndNodes", legacy ? 7 : 9)) return vstrdup("%s:%i: unexpected keyword", FileName, LineNum);
  if (!fgets(Line, 100, MeshFile)) return vstrdup("%s: bad file format (elements section not initiated)", FileName);
  LineNum++;
  if (strncmp(Line, legacy ? "$ELM" : "$Elements", legacy ? 4 : 9)) return vstrdup("%s:%i: unexpected keyword", FileName, LineNum);
  if (!fgets(Line, 100, MeshFile)) return vstrdup("%s: bad file format (invalid number of elements)", FileName);
  LineNum++;
  int NumElements;
  if (1 != sscanf(Line, "%i", &NumElements) || NumElements < 0) return vstrdup("%s:%i: invalid number of elements", FileName, LineNum);
  PVI.clear();
  for (int ne = 0; ne < NumElements; ne++)
   { if (!fgets(Line, 100, MeshFile)) return strdup("too few elements in input file");
     LineNum++;
     int VI[3] = { 0, 0, 0 }, ElType, RegPhys = 0;
     if (legacy)
      { int ElNum, RegElem, NodeCnt;
        int nConv = sscanf(Line, "%i %i %i %i %i %i %i %i", &ElNum, &ElType, &RegPhys, &RegElem, &NodeCnt, VI, VI + 1, VI + 2);
        if (nConv < 5) return vstrdup("%s:%i: invalid element specification", FileName, LineNum);
      }
     else
      { int ElNum, nTags, nRead = 0;
        int nConv = sscanf(Line, "%i %i %i%n", &ElNum, &ElType, &nTags, &nRead);
        int bufPos = nRead;
        if (nConv < 3) return vstrdup("%s:%i: invalid element specification", FileName, LineNum);
        if (nTags == 0) RegPhys = 0; else { sscanf(Line + bufPos, "%i%n", &RegPhys, &nRead); }
        bufPos += nRead;
        int iDummy;
        for (int nt = 0; nt < nTags - 1; nt++) { sscanf(Line + bufPos, "%i%n", &iDummy, &nRead); bufPos += nRead; }
        nConv = sscanf(Line + bufPos, "%i %i %i", VI, VI + 1, VI + 2);
        if (ElType == 2 && nConv != 3) return vstrdup("%s:%i: invalid element specification", FileName, LineNum);
      }
     if (ElType == 2 && (MeshTag == -1 || MeshTag == RegPhys))
      for (int i = 0; i < 3; i++)
       { if (VI[i] < 0 || VI[i] >= (int)GMSH2HR.size() || GMSH2HR[VI[i]] < 0) return vstrdup("%s:%i: invalid vertex index", FileName, LineNum);
         PVI.push_back(GMSH2HR[VI[i]]); }
   }
  fclose(MeshFile);
  return 0;
}

static char *ParseMeshFile(const char *MeshFileName, int MeshTag, std::vector<double> &VC, std::vector<int> &PVI)
{ FILE *f = OpenWithPath(g_meshDirs, MeshFileName);
  if (!f) return vstrdup("could not open file %s", MeshFileName);
  const char *ext = strrchr(MeshFileName, '.');
  if (!ext) return vstrdup("file %s: invalid extension", MeshFileName);
  if (!strcasecmp(ext + 1, "msh")) return ParseGMSHFile(f, MeshFileName, MeshTag, VC, PVI);
  return vstrdup("file %s: unknown extension %s", MeshFileName, ext + 1);
}

// ---------------------------------------------------------------------------------------
// RWGSurface
// ---------------------------------------------------------------------------------------
static void InitRWGPanel(RWGPanel *P, double *Vertices)
{ double *V[3] = { Vertices + 3 * P->VI[0], Vertices + 3 * P->VI[1], Vertices + 3 * P->VI[2] };
  for (int i = 0; i < 3; i++) P->Centroid[i] = (V[0][i] + V[1][i] + V[2][i]) / 3.0;
  P->Radius = VecDistance(P->Centroid, V[0]);
  P->Radius = fmax(P->Radius, VecDistance(P->Centroid, V[1]));
  P->Radius = fmax(P->Radius, VecDistance(P->Centroid, V[2]));
  double VA[3], VB[3];
  for (int i = 0; i < 3; i++) { VA[i] = V[1][i] - V[0][i]; VB[i] = V[2][i] - V[0][i]; }
  P->ZHat[0] = VA[1] * VB[2] - VA[2] * VB[1]; P->ZHat[1] = VA[2] * VB[0] - VA[0] * VB[2]; P->ZHat[2] = VA[0] * VB[1] - VA[1] * VB[0];
  if (P->ZHatFlipped) for (int i = 0; i < 3; i++) P->ZHat[i] *= -1.0;
  double n = sqrt(P->ZHat[0] * P->ZHat[0] + P->ZHat[1] * P->ZHat[1] + P->ZHat[2] * P->ZHat[2]);
  double in = 1.0 / n; for (int i = 0; i < 3; i++) P->ZHat[i] *= in;
  P->Area = 0.5 * n;
}

static char *ParseOBJECTSURFACESection(RWGSurface *S, FILE *f, int *LineNum)
{ char Line[MAXSTR];
  while (fgets(Line, MAXSTR, f))
   { (*LineNum)++;
     char LineCopy[MAXSTR]; strncpy(LineCopy, Line, MAXSTR);
     char *Tokens[MAXTOK]; int NumTokens = Tokenize(LineCopy, Tokens, MAXTOK);
     if (NumTokens == 0 || Tokens[0][0] == '#') continue;
     if (!strcasecmp(Tokens[0], "MESHFILE"))
      { if (NumTokens != 2) return strdup("MESHFILE keyword requires one argument"); S->MeshFileName = strdup(Tokens[1]); }
     else if (!strcasecmp(Tokens[0], "MESHTAG"))
      { if (NumTokens != 2) return strdup("MESHTAG keyword requires one argument"); sscanf(Tokens[1], "%i", &S->MeshTag); }
     else if (!strcasecmp(Tokens[0], "MATERIAL"))
      { if (NumTokens != 2) return strdup("MATERIAL keyword requires one argument");
        if (!S->IsObject) return strdup("MATERIAL keyword may not be used in SURFACE...ENDSURFACE sections");
        S->MaterialName = strdup(Tokens[1]);
        S->RegionLabels[0] = strdup("EXTERIOR"); S->RegionLabels[1] = strdup(S->Label);
        S->IsPEC = !strcasecmp(S->MaterialName, "PEC");
      }
     else if (!strcasecmp(Tokens[0], "REGIONS"))
      { if (S->IsObject) return strdup("REGIONS keyword may not be used in OBJECT...ENDOBJECT sections");
        if (NumTokens < 2 || NumTokens > 3) return strdup("REGIONS keyword requires one or two arguments");
        S->IsPEC = (NumTokens == 2);
        S->RegionLabels[0] = strdup(Tokens[1]); S->RegionLabels[1] = strdup(S->IsPEC ? "PEC" : Tokens[2]);
      }
     else if (!strcasecmp(Tokens[0], "DISPLACED") || !strcasecmp(Tokens[0], "ROTATED"))
      { GTrans *NewGT = new GTrans; char *Err = 0; int Consumed = 0;
        NewGT->Parse(Tokens, NumTokens, &Err, &Consumed);
        if (Err) return Err;
        if (Consumed != NumTokens) return strdup("junk at end of line");
        if (S->OTGT == 0) S->OTGT = NewGT;
        else { ((GTrans *)S->OTGT)->Transform(NewGT); delete NewGT; }
      }
     else if (!strcasecmp(Tokens[0], "SURFACE_IMPEDANCE"))
      { // dimensionless surface impedance Zeta(w,x,y,z): the rest of the line is the expression (RWGSurface.cc:188-210)
        if (NumTokens < 2) return strdup("no argument specified for SURFACE_IMPEDANCE");
        char *z = Line; while (*z && strchr(" \t\n", *z)) z++;
        z += strlen("SURFACE_IMPEDANCE");
        size_t n = strlen(z); while (n > 0 && (z[n - 1] == '\n' || z[n - 1] == '\r')) z[--n] = 0;
        char *Err = 0;
        S->SurfaceZeta = CExprCreate(z, &Err);
        if (S->SurfaceZeta == 0 || Err) return vstrdup("invalid SURFACE_IMPEDANCE %s", z);
      }
     else if (!strcasecmp(Tokens[0], "ENDOBJECT") || !strcasecmp(Tokens[0], "ENDSURFACE")) return 0;
     else return vstrdup("unknown keyword %s in %s section", Tokens[0], S->IsObject ? "OBJECT" : "SURFACE");
   }
  return vstrdup("unexpected end of file (unterminated %s section)", S->IsObject ? "OBJECT" : "SURFACE");
}

static void InitSurfaceFields(RWGSurface *S)
{ S->NumVertices = S->NumPanels = S->NumEdges = S->NumHalfRWGEdges = S->NumExteriorEdges = S->NumBFs = 0;
  S->Vertices = 0; S->Panels = 0; S->Edges = 0; S->HalfRWGEdges = 0; S->IsClosed = 0;
  S->IsPEC = true; S->IsObject = true; S->Index = 0; S->RegionIndices[0] = 0; S->RegionIndices[1] = -1;
  S->RegionLabels[0] = S->RegionLabels[1] = 0; S->MaterialName = 0; S->MeshFileName = 0; S->MeshTag = -1;
  S->Label = 0; S->ErrMsg = 0; S->tolVecClose = 0.0; S->TotalStraddlers = 0; S->OTGT = S->GT = 0; S->SurfaceZeta = 0;
}

RWGSurface::RWGSurface(FILE *f, const char *pLabel, int *LineNum, char *Keyword)
{ InitSurfaceFields(this);
  Label = strdup(pLabel);
  IsObject = !strcasecmp(Keyword, "OBJECT");
  ErrMsg = ParseOBJECTSURFACESection(this, f, LineNum);
  if (ErrMsg) return;
  if (MeshFileName == 0) { ErrMsg = vstrdup("%s section must include a MESHFILE specification", Keyword); return; }
  std::vector<double> VC; std::vector<int> PVI;
  ErrMsg = ParseMeshFile(MeshFileName, MeshTag, VC, PVI);
  if (ErrMsg) return;
  if (RegionLabels[0] == 0) { IsPEC = true; RegionLabels[0] = strdup("EXTERIOR"); RegionLabels[1] = strdup("PEC"); }
  ErrMsg = InitRWGSurface(VC, PVI);
}

RWGSurface::RWGSurface(const char *MeshFile, int pMeshTag)
{ InitSurfaceFields(this);
  MeshFileName = strdup(MeshFile); MeshTag = pMeshTag; Label = strdup(MeshFile);
  std::vector<double> VC; std::vector<int> PVI;
  ErrMsg = ParseMeshFile(MeshFileName, MeshTag, VC, PVI);
  if (ErrMsg) return;
  ErrMsg = InitRWGSurface(VC, PVI);
}

RWGSurface::~RWGSurface()
{ for (int i = 0; i < NumPanels; i++) free(Panels[i]);
  for (int i = 0; i < NumEdges; i++) free(Edges[i]);
  for (int i = 0; i < NumHalfRWGEdges; i++) free(HalfRWGEdges[i]);
  free(Panels); free(Edges); free(HalfRWGEdges); free(Vertices);
  free(Label); free(MeshFileName); free(MaterialName); free(RegionLabels[0]); free(RegionLabels[1]);
  delete (GTrans *)OTGT; delete (GTrans *)GT;
}

static void UpdateBoundingBox(RWGSurface *S)
{ // NOTE the reference only scans the first three vertices (RWGSurface.cc:573-584); kept as is
  // because RMin/RMax feed the PBC interpolation-table extents.
  double *V = S->Vertices;
  for (int i = 0; i < 3; i++) S->RMax[i] = S->RMin[i] = V[i];
  for (int nv = 0; nv < 3 && nv < S->NumVertices; nv++)
   for (int i = 0; i < 3; i++)
    { S->RMax[i] = fmax(S->RMax[i], V[3 * nv + i]); S->RMin[i] = fmin(S->RMin[i], V[3 * nv + i]); }
}

char *RWGSurface::InitRWGSurface(std::vector<double> &VC, std::vector<int> &PVI)
{ NumVertices = VC.size() / 3;
  if (NumVertices == 0 || PVI.empty()) return vstrdup("%s: mesh has no vertices or no triangles", MeshFileName);
  Vertices = (double *)malloc(3 * NumVertices * sizeof(double));
  memcpy(Vertices, &VC[0], 3 * NumVertices * sizeof(double));
  Origin[0] = Origin[1] = Origin[2] = 0.0;
  if (OTGT) { ((GTrans *)OTGT)->Apply(Vertices, NumVertices); ((GTrans *)OTGT)->Apply(Origin, 1); }
  NumPanels = PVI.size() / 3;
  Panels = (RWGPanel **)malloc(NumPanels * sizeof(RWGPanel *));
  for (int np = 0; np < NumPanels; np++)
   { RWGPanel *P = (RWGPanel *)calloc(1, sizeof(RWGPanel));
     P->VI[0] = PVI[3 * np]; P->VI[1] = PVI[3 * np + 1]; P->VI[2] = PVI[3 * np + 2];
     P->EI[0] = P->EI[1] = P->EI[2] = 0;
     P->Index = np; P->ZHatFlipped = false;
     InitRWGPanel(P, Vertices);
     Panels[np] = P;
   }
  char *Err = InitEdgeList();
  if (Err) return Err;
  if (RWGGeometry::UseHRWGFunctions)
   return strdup("SCUFF_HALF_RWG (half-RWG basis functions) is not supported by the device assembly path");
  for (int ne = 0; ne < NumEdges; ne++)
   { RWGEdge *E = Edges[ne];
     Panels[E->iPPanel]->EI[E->PIndex] = ne;
     if (E->iMPanel != -1) Panels[E->iMPanel]->EI[E->MIndex] = ne; }
  for (int ne = 0; ne < NumHalfRWGEdges; ne++)
   { RWGEdge *E = HalfRWGEdges[ne]; Panels[E->iPPanel]->EI[E->PIndex] = -(ne + 1); }
  NumBFs = IsPEC ? NumEdges : 2 * NumEdges;
  IsClosed = (NumExteriorEdges == 0);
  UpdateBoundingBox(this);
  return 0;
}

// Edge list: the order in which an edge is met the SECOND time while scanning panels 0..P-1,
// panel edges 0..2, is its index (and hence its BEM-matrix row).  InitEdgeList.cc:129-204.
char *RWGSurface::InitEdgeList()
{ std::vector<RWGEdge *> EdgeLists(NumVertices, (RWGEdge *)0);
  std::vector<RWGEdge *> interior;
  int NumTotalEdges = 0; NumEdges = 0;
  for (int np = 0; np < NumPanels; np++)
   for (int ne = 0; ne < 3; ne++)
    { RWGPanel *P = Panels[np];
      int iVLesser = P->VI[ne], iVGreater = P->VI[(ne + 1) % 3];
      if (iVLesser > iVGreater) std::swap(iVLesser, iVGreater);
      RWGEdge *E;
      for (E = EdgeLists[iVLesser]; E; E = E->Next) if (E->iV2 == iVGreater) break;
      if (E)
       { if (E->iMPanel != -1)
          return vstrdup("%s: invalid mesh topology: edge %i of panel %i also belongs to panels %i and %i ", MeshFileName, ne, np, E->iPPanel, E->iMPanel);
         E->iQM = P->VI[(ne + 2) % 3]; E->iMPanel = P->Index; E->MIndex = (ne + 2) % 3;
         E->Radius = fmax(E->Radius, VecDistance(E->Centroid, Vertices + 3 * E->iQM));
         E->Index = NumEdges++;
         interior.push_back(E);
       }
      else
       { NumTotalEdges++;
         E = (RWGEdge *)calloc(1, sizeof(RWGEdge));
         E->Next = EdgeLists[iVLesser]; EdgeLists[iVLesser] = E;
         E->iV1 = iVLesser; E->iV2 = iVGreater; E->iQP = P->VI[(ne + 2) % 3]; E->iQM = -1;
         double *VL = Vertices + 3 * iVLesser, *VG = Vertices + 3 * iVGreater;
         for (int i = 0; i < 3; i++) E->Centroid[i] = (VL[i] + VG[i]) / 2.0;
         E->Length = VecDistance(VL, VG);
         E->Radius = VecDistance(E->Centroid, Vertices + 3 * E->iQP);
         E->Radius = fmax(E->Radius, VecDistance(E->Centroid, VL));
         E->Radius = fmax(E->Radius, VecDistance(E->Centroid, VG));
         E->iPPanel = P->Index; E->PIndex = (ne + 2) % 3;
         E->iMPanel = -1; E->Index = -1; E->MIndex = (ne + 2) % 3;
       }
    }
  Edges = (RWGEdge **)malloc((NumEdges ? NumEdges : 1) * sizeof(RWGEdge *));
  HalfRWGEdges = (RWGEdge **)malloc((NumTotalEdges - NumEdges + 1) * sizeof(RWGEdge *));
  NumExteriorEdges = 0;
  for (int nv = 0; nv < NumVertices; nv++)
   for (RWGEdge *E = EdgeLists[nv]; E; E = E->Next)
    { if (E->Index == -1) { HalfRWGEdges[NumExteriorEdges] = E; E->Index = -(NumExteriorEdges + 1); NumExteriorEdges++; }
      else Edges[E->Index] = E; }
  NumHalfRWGEdges = NumExteriorEdges;
  for (int ne = 0; ne < NumEdges; ne++) Edges[ne]->Next = 0;
  for (int ne = 0; ne < NumHalfRWGEdges; ne++) HalfRWGEdges[ne]->Next = 0;
  return 0;
}

void RWGSurface::UpdatePanels()
{ for (int np = 0; np < NumPanels; np++) InitRWGPanel(Panels[np], Vertices);
  UpdateBoundingBox(this);
}

static void ApplyGT(RWGSurface *S, const GTrans *D)
{ D->Apply(S->Vertices, S->NumVertices);
  for (int ne = 0; ne < S->NumEdges; ne++) D->Apply(S->Edges[ne]->Centroid, 1);
  for (int ne = 0; ne < S->NumHalfRWGEdges; ne++) D->Apply(S->HalfRWGEdges[ne]->Centroid, 1);
  D->Apply(S->Origin, 1);
  S->UpdatePanels();
  if (!S->GT) S->GT = new GTrans(*D); else ((GTrans *)S->GT)->Transform(D);
}

void RWGSurface::Transform(const char *format, ...)
{ char buffer[MAXSTR]; va_list ap; va_start(ap, format); vsnprintf(buffer, MAXSTR, format, ap); va_end(ap);
  char *Tokens[MAXTOK]; int NumTokens = Tokenize(buffer, Tokens, MAXTOK);
  GTrans D; char *Err = 0; int pos = 0;
  while (pos < NumTokens)
   { GTrans One; int used = 0; One.Parse(Tokens + pos, NumTokens - pos, &Err, &used);
     if (Err) ErrExit("%s", Err);
     D.Transform(&One); pos += used; }
  ApplyGT(this, &D);
}

void RWGSurface::UnTransform()
{ if (!GT) return;
  GTrans *G = (GTrans *)GT;
  G->UnApply(Vertices, NumVertices);
  for (int ne = 0; ne < NumEdges; ne++) G->UnApply(Edges[ne]->Centroid, 1);
  for (int ne = 0; ne < NumHalfRWGEdges; ne++) G->UnApply(HalfRWGEdges[ne]->Centroid, 1);
  G->UnApply(Origin, 1);
  G->Reset();
  UpdatePanels();
}

// point-in-closed-surface by counting piercings of a plumb line dropped to z = -infinity
// (same parity rule as the reference's kd-tree version, PointInObject.cc:561-583)
bool RWGSurface::Contains(const double X0in[3])
{ if (!IsClosed) return false;
  // bounding-box rejection first (PointInObject.cc:574-580)
  double bmin[3], bmax[3];
  for (int i = 0; i < 3; i++) bmin[i] = bmax[i] = Vertices[3 * Panels[0]->VI[0] + i];
  for (int np = 0; np < NumPanels; np++)
   for (int j = 0; j < 3; j++)
    for (int i = 0; i < 3; i++)
     { double v = Vertices[3 * Panels[np]->VI[j] + i]; bmin[i] = fmin(bmin[i], v); bmax[i] = fmax(bmax[i], v); }
  for (int i = 0; i < 3; i++) if (X0in[i] < bmin[i] || X0in[i] > bmax[i]) return false;
  // nudge the plumb line off mesh vertices / edges (meshes displaced along z share x,y coordinates)
  double X0[3] = { X0in[0] + 1.2345678e-7 * (bmax[0] - bmin[0]), X0in[1] + 2.3456789e-7 * (bmax[1] - bmin[1]), X0in[2] };
  int piercings = 0;
  for (int np = 0; np < NumPanels; np++)
   { const double *A = Vertices + 3 * Panels[np]->VI[0], *B = Vertices + 3 * Panels[np]->VI[1], *C = Vertices + 3 * Panels[np]->VI[2];
     double d = (B[1] - C[1]) * (A[0] - C[0]) + (C[0] - B[0]) * (A[1] - C[1]);
     if (d == 0.0) continue;
     double l1 = ((B[1] - C[1]) * (X0[0] - C[0]) + (C[0] - B[0]) * (X0[1] - C[1])) / d;
     double l2 = ((C[1] - A[1]) * (X0[0] - C[0]) + (A[0] - C[0]) * (X0[1] - C[1])) / d;
     double l3 = 1.0 - l1 - l2;
     if (l1 < 0.0 || l2 < 0.0 || l3 < 0.0) continue;
     double z = l1 * A[2] + l2 * B[2] + l3 * C[2];
     if (z < X0[2]) piercings++;
   }
  return (piercings % 2) == 1;
}
bool RWGSurface::Contains(RWGSurface *S)
{ if (!IsClosed) return false;
  if (!S || S->NumPanels <= 0) return false;
  if (S->IsPEC) return false;
  return Contains(S->Vertices + 3 * S->Panels[0]->VI[0]);
}

// ---------------------------------------------------------------------------------------
// RWGGeometry
// ---------------------------------------------------------------------------------------
void RWGGeometry::AddRegion(const char *RegionLabel, const char *MaterialName, int LineNum)
{ int RegionID = 0;
  if (!strcasecmp(RegionLabel, "EXTERIOR")) RegionID = 0;
  else if (!strcasecmp(RegionLabel, "PEC")) ErrExit("%s:%i: REGIONs cannot be PEC (use an OBJECT instead)", GeoFileName, LineNum);
  else
   { RegionLabels = (char **)realloc(RegionLabels, (NumRegions + 1) * sizeof(char *));
     RegionLabels[NumRegions] = strdup(RegionLabel);
     RegionMPs = (MatProp **)realloc(RegionMPs, (NumRegions + 1) * sizeof(MatProp *));
     RegionMPs[NumRegions] = 0;
     NumRegions++; RegionID = NumRegions - 1;
   }
  MatProp *MP = new MatProp(MaterialName);
  if (MP->ErrMsg) ErrExit("%s:%i: %s", GeoFileName, LineNum, MP->ErrMsg);
  if (MP->IsPEC()) ErrExit("%s:%i: REGIONs cannot be PEC (use an OBJECT instead)", GeoFileName, LineNum);
  RegionMPs[RegionID] = MP;
}

int RWGGeometry::GetRegionByLabel(const char *Label)
{ if (!Label) return -1;
  for (int nr = 0; nr < NumRegions; nr++) if (!strcasecmp(Label, RegionLabels[nr])) return nr;
  return -1;
}
RWGSurface *RWGGeometry::GetSurfaceByLabel(const char *Label, int *pns)
{ if (pns) *pns = -1;
  if (!Label) return 0;
  for (int ns = 0; ns < NumSurfaces; ns++)
   if (!strcasecmp(Label, Surfaces[ns]->Label)) { if (pns) *pns = ns; return Surfaces[ns]; }
  return 0;
}

RWGGeometry::RWGGeometry(const char *pGeoFileName, int pLogLevel)
{ MatProp::SetLengthUnit(1.0e-6);
  LogLevel = pLogLevel;
  NumSurfaces = TotalBFs = TotalEdges = TotalPanels = 0;
  GeoFileName = strdup(pGeoFileName);
  Surfaces = 0; AllSurfacesClosed = 1;
  LDim = 0; LBasis = 0;
  for (int nd = 0; nd < MAXLDIM; nd++) { RegionIsExtended[nd] = 0; NumStraddlers[nd] = 0; }
  tolVecClose = 0.0;
  NumRegions = 1;
  RegionLabels = (char **)malloc(sizeof(char *)); RegionLabels[0] = strdup("EXTERIOR");
  RegionMPs = (MatProp **)malloc(sizeof(MatProp *)); RegionMPs[0] = new MatProp("VACUUM");
  for (int i = 0; i < 5; i++) LastPPICounts[i] = 0;

  const char *s;
  if ((s = getenv("SCUFF_LOGLEVEL")))
   { if (!strcasecmp(s, "NONE")) LogLevel = SCUFF_NOLOGGING; else if (!strcasecmp(s, "TERSE")) LogLevel = SCUFF_TERSELOGGING;
     else if (!strcasecmp(s, "VERBOSE")) LogLevel = SCUFF_VERBOSELOGGING; else if (!strcasecmp(s, "VERBOSE2")) LogLevel = SCUFF_VERBOSE2; }
  s = getenv("SCUFF_DISABLE_CACHE"); RWGGeometry::DisableCache = (s && s[0] == '1');
  s = getenv("SCUFF_HALF_RWG"); RWGGeometry::UseHRWGFunctions = (s && s[0] == '1');
  s = getenv("SCUFF_MATRIX_2018");
  if (s && s[0] == '1') ErrExit("SCUFF_MATRIX_2018 (the alternative 2018 assembler) is not supported by the device assembly path");

  FILE *f = fopen(GeoFileName, "r");
  std::string GeoDir = ".";
  if (!f)
   { const char *gp = getenv("SCUFF_GEO_PATH");
     if (gp) { std::string p = std::string(gp) + "/" + GeoFileName; f = fopen(p.c_str(), "r"); if (f) GeoDir = gp; } }
  else
   { const char *slash = strrchr(GeoFileName, '/'); if (slash) GeoDir = std::string(GeoFileName, slash - GeoFileName); }
  if (!f) ErrExit("could not open %s", GeoFileName);
  g_meshDirs.clear(); g_meshDirs.push_back(GeoDir); g_meshDirs.push_back(GeoDir + "/../mshFiles");

  char Line[MAXSTR]; int LineNum = 0;
  while (fgets(Line, MAXSTR, f))
   { LineNum++;
     char *Tokens[MAXTOK]; int nTokens = Tokenize(Line, Tokens, MAXTOK);
     if (nTokens == 0 || Tokens[0][0] == '#') continue;
     if (!strcasecmp(Tokens[0], "MEDIUM"))
      { char L2[MAXSTR]; bool done = false;
        while (!done && fgets(L2, MAXSTR, f))
         { LineNum++; char *T2[MAXTOK]; int n2 = Tokenize(L2, T2, MAXTOK);
           if (n2 == 0 || T2[0][0] == '#') continue;
           if (!strcasecmp(T2[0], "MATERIAL")) { if (n2 != 2) ErrExit("%s:%i: syntax error", GeoFileName, LineNum); AddRegion("EXTERIOR", T2[1], LineNum); }
           else if (!strcasecmp(T2[0], "ENDMEDIUM")) done = true;
           else ErrExit("%s:%i: unknown keyword %s", GeoFileName, LineNum, T2[0]); }
        if (!done) ErrExit("%s: unexpected end of file", GeoFileName);
      }
     else if (!strcasecmp(Tokens[0], "LATTICE"))
      { double LBV[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } }; char L2[MAXSTR]; bool done = false; LDim = 0;
        while (!done && fgets(L2, MAXSTR, f))
         { LineNum++; char *T2[MAXTOK]; int n2 = Tokenize(L2, T2, MAXTOK);
           if (n2 == 0 || T2[0][0] == '#') continue;
           if (!strcasecmp(T2[0], "VECTOR"))
            { if (LDim == MAXLDIM) ErrExit("%s:%i: too many lattice vectors", GeoFileName, LineNum);
              if (n2 < 2 || n2 > 4) ErrExit("%s:%i: lattice vectors must have 1, 2, or 3 components", GeoFileName, LineNum);
              for (int nc = 0; nc < n2 - 1; nc++)
               if (1 != sscanf(T2[nc + 1], "%le", &LBV[LDim][nc])) ErrExit("%s:%i: invalid vector component %s", GeoFileName, LineNum, T2[nc + 1]);
              LDim++; }
           else if (!strcasecmp(T2[0], "ENDLATTICE"))
            { LBasis = new HMatrix(3, LDim);
              for (int nd = 0; nd < LDim; nd++) for (int i = 0; i < 3; i++) LBasis->SetEntry(i, nd, LBV[nd][i]);
              done = true; }
           else ErrExit("%s:%i: unknown keyword %s", GeoFileName, LineNum, T2[0]); }
        if (!done) ErrExit("%s: unexpected end of file", GeoFileName);
      }
     else if (!strcasecmp(Tokens[0], "MESHPATH") || !strcasecmp(Tokens[0], "MESH_PATH"))
      { if (nTokens != 2) ErrExit("%s:%i: invalid MESHPATH specification", GeoFileName, LineNum);
        if (Tokens[1][0] == '/') g_meshDirs.push_back(Tokens[1]); else g_meshDirs.push_back(GeoDir + "/" + Tokens[1]); }
     else if (!strcasecmp(Tokens[0], "MATERIAL"))
      { if (nTokens == 1) ErrExit("%s:%i: no name given for MATERIAL ", GeoFileName, LineNum);
        else if (nTokens > 2) ErrExit("%s:%i: syntax error", GeoFileName, LineNum);
        std::string name = Tokens[1];
        char *Err = AddMaterialToMatPropDataBase(f, GeoFileName, name.c_str(), &LineNum);
        if (Err) ErrExit("%s:%i: %s", GeoFileName, LineNum, Err); }
     else if (!strcasecmp(Tokens[0], "REGION"))
      { if (nTokens != 4 || strcasecmp(Tokens[2], "MATERIAL")) ErrExit("%s:%i: syntax error", GeoFileName, LineNum);
        AddRegion(Tokens[1], Tokens[3], LineNum); }
     else if (!strcasecmp(Tokens[0], "OBJECT") || !strcasecmp(Tokens[0], "SURFACE"))
      { RWGSurface *S = 0; char Keyword[16]; strncpy(Keyword, Tokens[0], 15); Keyword[15] = 0;
        if (nTokens == 2) { std::string lbl = Tokens[1]; S = new RWGSurface(f, lbl.c_str(), &LineNum, Keyword); }
        else if (nTokens != 1) ErrExit("%s:%i: syntax error", GeoFileName, LineNum);
        else { char Label[MAXSTR]; snprintf(Label, MAXSTR, "Surface_%i", NumSurfaces + 1); S = new RWGSurface(f, Label, &LineNum, Keyword); }
        if (S->ErrMsg) ErrExit("%s:%i: %s", GeoFileName, LineNum, S->ErrMsg);
        if (S->IsObject)
         { S->RegionIndices[0] = 0;
           if (S->IsPEC) S->RegionIndices[1] = -1;
           else { AddRegion(S->RegionLabels[1], S->MaterialName, LineNum); S->RegionIndices[1] = NumRegions - 1; } }
        else
         { AllSurfacesClosed = 0;
           S->RegionIndices[0] = GetRegionByLabel(S->RegionLabels[0]);
           if (S->RegionIndices[0] == -1) ErrExit("%s:%i: unknown region %s", GeoFileName, LineNum, S->RegionLabels[0]);
           if (S->IsPEC) S->RegionIndices[1] = -1;
           else { S->RegionIndices[1] = GetRegionByLabel(S->RegionLabels[1]);
                  if (S->RegionIndices[1] == -1) ErrExit("%s:%i: unknown region %s", GeoFileName, LineNum, S->RegionLabels[1]); } }
        NumSurfaces++;
        Surfaces = (RWGSurface **)realloc(Surfaces, NumSurfaces * sizeof(RWGSurface *));
        Surfaces[NumSurfaces - 1] = S; S->Index = NumSurfaces - 1;
      }
     else if (!strcasecmp(Tokens[0], "SUBSTRATE")) ErrExit("%s:%i: SUBSTRATE is not supported by the device assembly path", GeoFileName, LineNum);
     else ErrExit("%s:%i: syntax error", GeoFileName, LineNum);
   }
  fclose(f);
  if (NumSurfaces == 0) ErrExit("%s: geometry contains no surfaces", GeoFileName);

  tolVecClose = 1.0 / 0.0;
  for (int ns = 0; ns < NumSurfaces; ns++)
   for (int ne = 0; ne < Surfaces[ns]->NumEdges; ne++) tolVecClose = fmin(tolVecClose, Surfaces[ns]->Edges[ne]->Length);
  tolVecClose *= 1e-3;
  for (int ns = 0; ns < NumSurfaces; ns++) Surfaces[ns]->tolVecClose = tolVecClose;

  if (LBasis) { LDim = LBasis->NC; InitPBCData(); }

  // nesting autodetection + topological sort (RWGGeometry.cc:583-649)
  for (int ns = 1; ns < NumSurfaces; ++ns)
   { RWGSurface *S = Surfaces[ns]; int NewSlotForS = -1;
     for (int nsp = ns - 1; nsp >= 0; --nsp)
      { RWGSurface *SP = Surfaces[nsp];
        if (SP->Contains(S))
         { if (SP->IsPEC) ErrExit("%s: PEC object %s cannot contain object %s", GeoFileName, SP->Label, S->Label);
           S->RegionIndices[0] = SP->RegionIndices[1];
           free(S->RegionLabels[0]); S->RegionLabels[0] = strdup(SP->RegionLabels[1]);
           break; }
        else if (S->Contains(SP)) NewSlotForS = nsp;
      }
     if (NewSlotForS != -1)
      { for (int nsp = ns; nsp > NewSlotForS; --nsp)
         { Surfaces[nsp] = Surfaces[nsp - 1]; RWGSurface *SP = Surfaces[nsp];
           if (SP->RegionIndices[0] == S->RegionIndices[0] && S->Contains(SP))
            { if (S->IsPEC) ErrExit("%s: PEC object %s cannot contain object %s", GeoFileName, S->Label, SP->Label);
              SP->RegionIndices[0] = S->RegionIndices[1];
              free(SP->RegionLabels[0]); SP->RegionLabels[0] = strdup(S->RegionLabels[1]); } }
        Surfaces[NewSlotForS] = S; }
   }
  for (int ns = 0; ns < NumSurfaces; ns++) Surfaces[ns]->Index = ns;
  // (panel normals are not re-oriented here: they do not enter the BEM matrix)

  BFIndexOffset = (int *)malloc(NumSurfaces * sizeof(int));
  EdgeIndexOffset = (int *)malloc(NumSurfaces * sizeof(int));
  PanelIndexOffset = (int *)malloc(NumSurfaces * sizeof(int));
  TotalBFs = Surfaces[0]->NumBFs; TotalEdges = Surfaces[0]->NumEdges; TotalPanels = Surfaces[0]->NumPanels;
  BFIndexOffset[0] = EdgeIndexOffset[0] = PanelIndexOffset[0] = 0;
  for (int ns = 1; ns < NumSurfaces; ns++)
   { TotalBFs += Surfaces[ns]->NumBFs; TotalEdges += Surfaces[ns]->NumEdges; TotalPanels += Surfaces[ns]->NumPanels;
     BFIndexOffset[ns] = BFIndexOffset[ns - 1] + Surfaces[ns - 1]->NumBFs;
     EdgeIndexOffset[ns] = EdgeIndexOffset[ns - 1] + Surfaces[ns - 1]->NumEdges;
     PanelIndexOffset[ns] = PanelIndexOffset[ns - 1] + Surfaces[ns - 1]->NumPanels; }
  StoredOmega = 0.0;
  EpsTF = (cdouble *)calloc(NumRegions, sizeof(cdouble));
  MuTF = (cdouble *)calloc(NumRegions, sizeof(cdouble));

  // Mate[]: same mesh file + tag and same materials on both sides (RWGGeometry.cc:738-768)
  Mate = (int *)malloc(NumSurfaces * sizeof(int)); Mate[0] = -1;
  for (int ns = 1; ns < NumSurfaces; ns++)
   { RWGSurface *S = Surfaces[ns]; Mate[ns] = -1;
     int nr1 = S->RegionIndices[0], nr2 = S->RegionIndices[1];
     for (int nsp = 0; nsp < ns && Mate[ns] == -1; nsp++)
      { RWGSurface *SP = Surfaces[nsp]; int nr1p = SP->RegionIndices[0], nr2p = SP->RegionIndices[1];
        if (!strcmp(S->MeshFileName, SP->MeshFileName) && S->MeshTag == SP->MeshTag
            && !strcmp(RegionMPs[nr1]->Name, RegionMPs[nr1p]->Name)
            && ((S->IsPEC && SP->IsPEC) || (!S->IsPEC && !SP->IsPEC && !strcmp(RegionMPs[nr2]->Name, RegionMPs[nr2p]->Name))))
         Mate[ns] = nsp; } }
  SurfaceMoved = (int *)calloc(NumSurfaces, sizeof(int));
  DeviceDirty.assign(NumSurfaces, 1);
  DeviceId = 0;
}

RWGGeometry::~RWGGeometry()
{ for (int ns = 0; ns < NumSurfaces; ns++) delete Surfaces[ns];
  free(Surfaces);
  for (int nr = 0; nr < NumRegions; nr++) { free(RegionLabels[nr]); delete RegionMPs[nr]; }
  free(RegionLabels); free(RegionMPs);
  free(BFIndexOffset); free(EdgeIndexOffset); free(PanelIndexOffset); free(EpsTF); free(MuTF); free(Mate); free(SurfaceMoved);
  for (int nd = 0; nd < MAXLDIM; nd++) { free(RegionIsExtended[nd]); free(NumStraddlers[nd]); }
  delete LBasis; free(GeoFileName);
}

void RWGGeometry::UpdateCachedEpsMuValues(cdouble Omega)
{ StoredOmega = Omega;
  for (int nr = 0; nr < NumRegions; nr++) RegionMPs[nr]->GetEpsMu(Omega, &EpsTF[nr], &MuTF[nr]);
}

// "TRANS name SURFACE label DISP x y z ROT t ABOUT x y z SURFACE label2 ..."
void RWGGeometry::Transform(const char *TransLine)
{ if (strlen(TransLine) >= MAXSTR) ErrExit("Transform: line longer than %i characters", MAXSTR - 1);
  char buf[MAXSTR]; strncpy(buf, TransLine, MAXSTR - 1); buf[MAXSTR - 1] = 0;
  char *Tokens[MAXTOK]; int n = Tokenize(buf, Tokens, MAXTOK);
  int pos = 0;
  if (n >= 2 && (!strcasecmp(Tokens[0], "TRANS") || !strcasecmp(Tokens[0], "TRANSFORMATION"))) pos = 2;
  memset(SurfaceMoved, 0, NumSurfaces * sizeof(int));
  while (pos < n)
   { if (strcasecmp(Tokens[pos], "SURFACE") && strcasecmp(Tokens[pos], "OBJECT")) ErrExit("Transform: syntax error at %s", Tokens[pos]);
     if (pos + 1 >= n) ErrExit("Transform: missing surface label");
     int ns; RWGSurface *S = GetSurfaceByLabel(Tokens[pos + 1], &ns);
     if (!S) ErrExit("Transform: unknown surface %s", Tokens[pos + 1]);
     pos += 2; GTrans D;
     while (pos < n && strcasecmp(Tokens[pos], "SURFACE") && strcasecmp(Tokens[pos], "OBJECT"))
      { GTrans One; char *Err = 0; int used = 0; One.Parse(Tokens + pos, n - pos, &Err, &used);
        if (Err) ErrExit("Transform: %s", Err);
        D.Transform(&One); pos += used; }
     ApplyGT(S, &D);
     SurfaceMoved[ns] = 1; DeviceDirty[ns] = 1;
   }
}
void RWGGeometry::UnTransform()
{ for (int ns = 0; ns < NumSurfaces; ns++)
   if (Surfaces[ns]->GT) { Surfaces[ns]->UnTransform(); DeviceDirty[ns] = 1; }
}

// ---------------------------------------------------------------------------------------------------------------
// Periodic geometries: straddlers (reference semantics: PBCSetup.cc:64-118 FindPartnerEdge, :142-302 AddStraddlers,
// :307-372 InitPBCData).  An exterior edge on the unit-cell boundary whose translate by a lattice vector is another
// exterior edge becomes a full RWG edge: one vertex (the partner's Q vertex shifted back by the lattice vector) and one
// panel {iV1, iV2, new vertex} are appended, and the edge is appended to Edges[] AFTER the interior edges — which fixes
// the matrix row/column order of periodic geometries.
// ---------------------------------------------------------------------------------------------------------------
static bool VecCloseL1(const double *a, const double *b, double abstol)      // libhrutil/Vector.cc:294-297
{ return fabs(a[0] - b[0]) + fabs(a[1] - b[1]) + fabs(a[2] - b[2]) <= 3 * abstol; }

void RWGSurface::AddStraddlers(HMatrix *LBasis, int NumStraddlersOut[MAXLDIM])
{ const int LDim = LBasis->NC;
  for (int nd = 0; nd < MAXLDIM; nd++) NumStraddlersOut[nd] = 0;
  std::vector<double> NewV; std::vector<RWGPanel *> NewP; std::vector<RWGEdge *> NewE;
  for (int nei = 0; nei < NumExteriorEdges; nei++)
   { RWGEdge *E = HalfRWGEdges[nei];
     if (!E) continue;
     const double *V1 = Vertices + 3 * E->iV1, *V2 = Vertices + 3 * E->iV2;
     int partner = -1, whichBV = 0; double V[3];
     for (int nd = 0; nd < LDim && partner < 0; nd++)
      { double V1T[3], V2T[3];
        for (int nc = 0; nc < 3; nc++) { V1T[nc] = V1[nc] + LBasis->GetEntryD(nc, nd); V2T[nc] = V2[nc] + LBasis->GetEntryD(nc, nd); }
        for (int neip = 0; neip < NumExteriorEdges; neip++)
         { RWGEdge *EP = HalfRWGEdges[neip];
           if (!EP) continue;
           const double *V1P = Vertices + 3 * EP->iV1, *V2P = Vertices + 3 * EP->iV2;
           if ((VecCloseL1(V1T, V1P, tolVecClose) && VecCloseL1(V2T, V2P, tolVecClose))
               || (VecCloseL1(V1T, V2P, tolVecClose) && VecCloseL1(V2T, V1P, tolVecClose)))
            { for (int nc = 0; nc < 3; nc++) V[nc] = Vertices[3 * EP->iQP + nc];
              for (int nc = 0; nc < 3; nc++) V[nc] -= LBasis->GetEntryD(nc, nd);
              NumStraddlersOut[nd]++; whichBV = nd; partner = neip; break; }
         }
      }
     (void)whichBV;
     if (partner < 0) continue;
     const int NumNew = (int)NewE.size();
     NewV.insert(NewV.end(), V, V + 3);
     for (int nc = 0; nc < 3; nc++) { RMax[nc] = fmax(RMax[nc], V[nc]); RMin[nc] = fmin(RMin[nc], V[nc]); }
     E->iQM = NumVertices + NumNew;
     E->iMPanel = NumPanels + NumNew;
     E->Index = NumEdges + NumNew;
     E->MIndex = 2;
     E->Radius = VecDistance(E->Centroid, Vertices + 3 * E->iQP);
     E->Radius = fmax(E->Radius, VecDistance(E->Centroid, V));
     E->Radius = fmax(E->Radius, VecDistance(E->Centroid, Vertices + 3 * E->iV1));
     E->Radius = fmax(E->Radius, VecDistance(E->Centroid, Vertices + 3 * E->iV2));
     NewE.push_back(E);
     HalfRWGEdges[nei] = 0;
     RWGPanel *P = (RWGPanel *)calloc(1, sizeof(RWGPanel));
     P->VI[0] = E->iV1; P->VI[1] = E->iV2; P->VI[2] = E->iQM;
     P->EI[0] = P->EI[1] = P->EI[2] = -1;
     P->Index = NumPanels + NumNew; P->ZHatFlipped = false;
     NewP.push_back(P);
   }
  const int NumNew = (int)NewE.size();
  TotalStraddlers = NumNew;
  if (NumNew == 0) return;
  Vertices = (double *)realloc(Vertices, 3 * (size_t)(NumVertices + NumNew) * sizeof(double));
  memcpy(Vertices + 3 * NumVertices, NewV.data(), 3 * (size_t)NumNew * sizeof(double));
  NumVertices += NumNew;
  Edges = (RWGEdge **)realloc(Edges, (size_t)(NumEdges + NumNew) * sizeof(RWGEdge *));
  memcpy(Edges + NumEdges, NewE.data(), NumNew * sizeof(RWGEdge *));
  NumEdges += NumNew;
  Panels = (RWGPanel **)realloc(Panels, (size_t)(NumPanels + NumNew) * sizeof(RWGPanel *));
  memcpy(Panels + NumPanels, NewP.data(), NumNew * sizeof(RWGPanel *));
  NumPanels += NumNew;
  for (int np = NumPanels - NumNew; np < NumPanels; np++) InitRWGPanel(Panels[np], Vertices);
  // defragment the list of exterior edges
  int keep = 0;
  for (int nei = 0; nei < NumExteriorEdges; nei++)
   if (HalfRWGEdges[nei]) { HalfRWGEdges[keep] = HalfRWGEdges[nei]; HalfRWGEdges[keep]->Index = keep; keep++; }
  NumExteriorEdges = NumHalfRWGEdges = keep;
  NumBFs = IsPEC ? NumEdges : 2 * NumEdges;
  for (int ne = 0; ne < NumEdges; ne++)
   { RWGEdge *E = Edges[ne];
     Panels[E->iPPanel]->EI[E->PIndex] = ne;
     if (E->iMPanel >= 0) Panels[E->iMPanel]->EI[E->MIndex] = ne; }
  for (int ne = 0; ne < NumHalfRWGEdges; ne++)
   { RWGEdge *E = HalfRWGEdges[ne]; Panels[E->iPPanel]->EI[E->PIndex] = -(ne + 1); }
}

void RWGGeometry::InitPBCData()
{ TotalBFs = TotalPanels = 0;
  for (int nd = 0; nd < LDim; nd++)
   { NumStraddlers[nd] = (int *)calloc(NumSurfaces, sizeof(int));
     RegionIsExtended[nd] = (bool *)calloc(NumRegions, sizeof(bool));
     RegionIsExtended[nd][0] = true; }     // the exterior medium is always extended
  for (int ns = 0; ns < NumSurfaces; ns++)
   { RWGSurface *S = Surfaces[ns];
     int NS[MAXLDIM];
     S->AddStraddlers(LBasis, NS);
     const int nr1 = S->RegionIndices[0], nr2 = S->RegionIndices[1];
     int Total = 0;
     for (int nd = 0; nd < LDim; nd++)
      { NumStraddlers[nd][ns] = NS[nd]; Total += NS[nd];
        if (NS[nd] > 0) { RegionIsExtended[nd][nr1] = true; if (nr2 != -1) RegionIsExtended[nd][nr2] = true; } }
     TotalBFs += S->NumBFs; TotalPanels += S->NumPanels;
     S->IsClosed = (Total == S->NumExteriorEdges);      // as written in the reference (PBCSetup.cc:364)
   }
}

void RWGGeometry::SyncDeviceGeometry()
{ scuffcuda_context *ctx = GetCudaContext();
  // the device context holds ONE geometry at a time; a process-unique id (not the object address,
  // which the allocator may recycle) tells whether it is ours
  static thread_local unsigned long long deviceOwner = 0;
  static std::atomic<unsigned long long> nextId(0);
  static thread_local scuffcuda_context *ownerCtx = 0;
  if (ownerCtx != ctx) { ownerCtx = ctx; deviceOwner = 0; }        // a new context (SetCudaDevice) holds no geometry yet
  if (DeviceId == 0) DeviceId = ++nextId;
  if (deviceOwner != DeviceId)
   { if (scuffcuda_set_num_surfaces(ctx, 0) || scuffcuda_set_num_surfaces(ctx, NumSurfaces)) ErrExit("%s", scuffcuda_last_error(ctx));
     DeviceDirty.assign(NumSurfaces, 1); deviceOwner = DeviceId; }
  for (int ns = 0; ns < NumSurfaces; ns++)
   { if (!DeviceDirty[ns]) continue;
     RWGSurface *S = Surfaces[ns];
     std::vector<int> pvi(3 * S->NumPanels), et(8 * (size_t)S->NumEdges);
     std::vector<double> pc(3 * S->NumPanels), pr(S->NumPanels), el(S->NumEdges);
     for (int np = 0; np < S->NumPanels; np++)
      { RWGPanel *P = S->Panels[np];
        for (int i = 0; i < 3; i++) { pvi[3 * np + i] = P->VI[i]; pc[3 * np + i] = P->Centroid[i]; }
        pr[np] = P->Radius; }
     for (int ne = 0; ne < S->NumEdges; ne++)
      { RWGEdge *E = S->Edges[ne]; int *e = &et[8 * (size_t)ne];
        e[0] = E->iV1; e[1] = E->iV2; e[2] = E->iQP; e[3] = E->iQM; e[4] = E->iPPanel; e[5] = E->iMPanel; e[6] = E->PIndex; e[7] = E->MIndex;
        el[ne] = E->Length; }
     if (scuffcuda_set_surface(ctx, ns, S->NumVertices, S->Vertices, S->NumPanels, pvi.data(), pc.data(), pr.data(),
                               S->NumEdges, et.data(), el.data(), S->IsPEC ? 1 : 0))
      ErrExit("%s", scuffcuda_last_error(ctx));
     DeviceDirty[ns] = 0;
   }
}

// \int f_alpha . f_beta over the panels the two RWG functions share (OPFT.cc:81-226, Overlaps[0]).  On a shared panel with
// source vertices Q_a, Q_b: f_a . f_b = +-(l_a l_b / 4A^2) (X - Q_a).(X - Q_b); with X - Q_a = s L1 + t L2 over the panel
// (L1 = Q_a' - Q_a, L2 = Q_a'' - Q_a') the integral is (l_a l_b / 2A) [ (L1.L1 + L1.L2)/4 + L1.DQ/3 + L2.L2/12 + L2.DQ/6 ].
static double PanelOverlap(RWGSurface *S, RWGPanel *P, int iQa, int iQb, double Sign, double LL)
{ const double *Qa = S->Vertices + 3 * P->VI[iQa], *Qa1 = S->Vertices + 3 * P->VI[(iQa + 1) % 3], *Qa2 = S->Vertices + 3 * P->VI[(iQa + 2) % 3];
  const double *Qb = S->Vertices + 3 * P->VI[iQb];
  double L1dL1 = 0, L1dL2 = 0, L1dDQ = 0, L2dL2 = 0, L2dDQ = 0;
  for (int i = 0; i < 3; i++)
   { const double l1 = Qa1[i] - Qa[i], l2 = Qa2[i] - Qa1[i], dq = Qa[i] - Qb[i];
     L1dL1 += l1 * l1; L1dL2 += l1 * l2; L1dDQ += l1 * dq; L2dL2 += l2 * l2; L2dDQ += l2 * dq; }
  return Sign * LL / (2.0 * P->Area) * ((L1dL1 + L1dL2) / 4.0 + L1dDQ / 3.0 + L2dL2 / 12.0 + L2dDQ / 6.0);
}
double RWGSurface::GetOverlap(int neAlpha, int neBeta)
{ RWGEdge *Ea = Edges[neAlpha], *Eb = Edges[neBeta];
  const double LL = Ea->Length * Eb->Length;
  double O = 0.0;
  if (Ea->iPPanel == Eb->iPPanel) O += PanelOverlap(this, Panels[Ea->iPPanel], Ea->PIndex, Eb->PIndex, 1.0, LL);
  if (Ea->iPPanel == Eb->iMPanel) O += PanelOverlap(this, Panels[Ea->iPPanel], Ea->PIndex, Eb->MIndex, -1.0, LL);
  if (Ea->iMPanel != -1 && Ea->iMPanel == Eb->iPPanel) O += PanelOverlap(this, Panels[Ea->iMPanel], Ea->MIndex, Eb->PIndex, -1.0, LL);
  if (Ea->iMPanel != -1 && Ea->iMPanel == Eb->iMPanel) O += PanelOverlap(this, Panels[Ea->iMPanel], Ea->MIndex, Eb->MIndex, 1.0, LL);
  return O;
}

}  // namespace scuff
