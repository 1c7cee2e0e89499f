// qjs_run: evaluates JavaScript files with Qt's own ECMAScript engine (QJSEngine, the "V4" engine of
// libQt6Qml) as shipped inside Nsight Compute's host directory of this image - the only JavaScript engine here.
// No Qt headers are installed, so the handful of entry points used are declared by their Itanium-ABI names
// and called with hand-written layouts (Qt 6.6: QString / QByteArray = {d, ptr, size}; QJSValue = one word;
// QObject-derived classes get generous opaque storage).  Test infrastructure only: it exists to run the
// UNMODIFIED reference sources (/root/reference/modem.js, app.js) and dump golden vectors (node/dump_golden.js).
//
// usage: qjs_run file1.js [file2.js ...]   - files are evaluated in order in one global scope; the string value
// of the global `__out` (or of the last expression) is written to stdout.  print(x) writes to stderr.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

struct QStr { void* d; char16_t* ptr; long long size; };          // QString, Qt 6
struct QBA  { void* d; char* ptr; long long size; };              // QByteArray, Qt 6
struct QBAView { long long size; const char* data; };             // QByteArrayView
struct QJSVal { unsigned long long d; };                            // QJSValue

extern "C" {
void qcoreapp_ctor(void* self, int& argc, char** argv, int flags) asm("_ZN16QCoreApplicationC1ERiPPci");
void qjsengine_ctor(void* self) asm("_ZN9QJSEngineC1Ev");
void qjsengine_evaluate(QJSVal* ret, void* self, const QStr* program, const QStr* file, int line, void* stack)
    asm("_ZN9QJSEngine8evaluateERK7QStringS2_iP5QListIS0_E");
void qjsengine_global(QJSVal* ret, const void* self) asm("_ZNK9QJSEngine12globalObjectEv");
void qstring_fromutf8(QStr* ret, QBAView v) asm("_ZN7QString8fromUtf8E14QByteArrayView");
void qstring_toutf8(QBA* ret, const QStr* s) asm("_ZN7QString13toUtf8_helperERKS_");
void qjsvalue_tostring(QStr* ret, const QJSVal* self) asm("_ZNK8QJSValue8toStringEv");
bool qjsvalue_iserror(const QJSVal* self) asm("_ZNK8QJSValue7isErrorEv");
bool qjsvalue_isundefined(const QJSVal* self) asm("_ZNK8QJSValue11isUndefinedEv");
void qjsvalue_property(QJSVal* ret, const QJSVal* self, const QStr* name) asm("_ZNK8QJSValue8propertyERK7QString");
}

static QStr mk(const std::string& s) { QStr q; qstring_fromutf8(&q, QBAView{(long long)s.size(), s.data()}); return q; }
static std::string str(const QStr& q) { QBA b; qstring_toutf8(&b, &q); return std::string(b.ptr ? b.ptr : "", (size_t)b.size); }

static std::string slurp(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "qjs_run: cannot open %s\n", path); exit(2); }
    std::string s; char buf[1 << 16]; size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, n);
    fclose(f);
    return s;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: qjs_run file.js [...]\n"); return 2; }
    alignas(16) static char app[256], eng[256];
    int ac = 1; char* av[2] = {argv[0], nullptr};
    qcoreapp_ctor(app, ac, av, 0x060603);
    qjsengine_ctor(eng);
    QJSVal last{0};
    for (int i = 1; i < argc; ++i) {
        QStr prog = mk(slurp(argv[i])), name = mk(argv[i]);
        qjsengine_evaluate(&last, eng, &prog, &name, 1, nullptr);
        if (qjsvalue_iserror(&last)) {
            QStr m; qjsvalue_tostring(&m, &last);
            QStr ln = mk("lineNumber"); QJSVal l; qjsvalue_property(&l, &last, &ln); QStr ls; qjsvalue_tostring(&ls, &l);
            QStr stn = mk("stack"); QJSVal st; qjsvalue_property(&st, &last, &stn); QStr sts; qjsvalue_tostring(&sts, &st);
            fprintf(stderr, "qjs_run: %s: uncaught %s (line %s)\n%s\n", argv[i], str(m).c_str(), str(ls).c_str(), str(sts).c_str());
            return 1;
        }
    }
    QJSVal g, out; qjsengine_global(&g, eng);
    QStr on = mk("__out"); qjsvalue_property(&out, &g, &on);
    const QJSVal* v = qjsvalue_isundefined(&out) ? &last : &out;
    QStr s; qjsvalue_tostring(&s, v);
    std::string u = str(s);
    fwrite(u.data(), 1, u.size(), stdout);
    fflush(stdout);
    _Exit(0);   // no teardown of the hand-laid-out Qt objects
}
