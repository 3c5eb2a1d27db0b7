// Error/log conventions and small helpers of the host-side mirror.
// Reference: libs/libhrutil/libhrutil.h:100-151 (ErrExit/Warn/Log), libhrutil.cc:228-235 (Tokenize).
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include "host_internal.h"

namespace scuff {

static ErrExitHandler g_handler = 0;
static std::string g_logFile;
// one device context per HOST THREAD (thread_local): the main thread's context serves ordinary single-GPU use; a SweepDriver
// worker thread calls SetCudaDevice(d) and gets its own context on device d.  Objects that hold device handles (HMatrix,
// RWGGeometry after its first assembly) belong to the thread that created their device twin.
static thread_local scuffcuda_context *g_ctx = 0;
static thread_local int g_device = -1;

void SetErrExitHandler(ErrExitHandler h) { g_handler = h; }
void SetLogFileName(const char *name) { g_logFile = name ? name : ""; }

void ErrExit(const char *fmt, ...)
{ char buf[2048]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  Log("error: %s (aborting)", buf);
  if (g_handler) { g_handler(buf); }
  fprintf(stderr, "error: %s (aborting)\n", buf);
  exit(1);
}
void Warn(const char *fmt, ...)
{ char buf[2048]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  fprintf(stderr, "**warning: %s \n", buf);
  Log("warning: %s", buf);
}
void Log(const char *fmt, ...)
{ if (g_logFile.empty()) return;
  FILE *f = fopen(g_logFile.c_str(), "a"); if (!f) return;
  char buf[2048]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  time_t t = time(0); struct tm tmv; localtime_r(&t, &tmv); char ts[64];
  strftime(ts, sizeof ts, "%D::%T", &tmv);
  fprintf(f, "%s: %s\n", ts, buf); fclose(f);
}

// number of live objects that hold handles into the context (device matrices, uploaded geometries)
static thread_local int g_liveDeviceObjects = 0;
void AddLiveDeviceObject(int delta) { g_liveDeviceObjects += delta; }

void SetCudaDevice(int device)
{ if (g_ctx && device != g_device)
   { // matrices / geometries hold handles into the old context: switching under them would leave them dangling
     if (g_liveDeviceObjects > 0)
      ErrExit("SetCudaDevice(%i): %i device-resident matrices/geometries are still alive on device %i; destroy them first", device, g_liveDeviceObjects, g_device);
     scuffcuda_destroy(g_ctx); g_ctx = 0; }
  g_device = device;
}

void ReleaseCudaContext()
{ if (g_ctx) { scuffcuda_destroy(g_ctx); g_ctx = 0; } }

scuffcuda_context *GetCudaContext()
{ if (g_ctx) return g_ctx;
  int dev = g_device;
  if (dev < 0)
   { const char *s = getenv("SCUFF_CUDA_DEVICE"); if (!s) s = getenv("LOCAL_RANK");
     dev = s ? atoi(s) : 0; }
  if (scuffcuda_create(dev, &g_ctx))
   ErrExit("libscuff_cuda: %s", scuffcuda_last_error(0));
  g_device = dev;
  return g_ctx;
}

// whitespace tokenizer (in place)
int Tokenize(char *s, char **Tokens, int MaxTokens, const char *Separators)
{ int n = 0; char *save = 0;
  for (char *t = strtok_r(s, Separators, &save); t; t = strtok_r(0, Separators, &save))
   { if (n >= MaxTokens) ErrExit("line has more than %i tokens (near '%s')", MaxTokens, t);     // never hand back more tokens than were stored
     Tokens[n++] = t; }
  return n;
}

char *vstrdup(const char *fmt, ...)
{ char buf[2048]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  return strdup(buf);
}

}  // namespace scuff
