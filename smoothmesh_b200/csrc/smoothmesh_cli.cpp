// smoothmesh_cli.cpp -- stand-alone `smoothMesh` executable with the reference's
// command-line surface (src/smoothMesh.C:1642-1784), polyMesh I/O and log
// lines, driving the GPU library through its C ABI only.
//
// Differences from the OpenFOAM build, all by necessity:
//  * options are parsed here instead of by argList; -case/-time/-parallel are
//    accepted like OpenFOAM's standard options;
//  * system/controlDict is read for deltaT, writeFormat, writePrecision and
//    startFrom/startTime only;
//  * features outside the hot path (boundary-layer treatment, boundary point
//    smoothing) are refused loudly when the options/files would enable them --
//    there is no CPU fallback.
#include "../../include/smgpu.h"
#include "../../include/smmesh.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dirent.h>
#include <fstream>
#include <condition_variable>
#include <ctime>
#include <map>
#include <mutex>
#include <thread>
#include <regex>
#include <sstream>
#include <string>
#include <sys/stat.h>
#include <vector>

static bool fileExists(const std::string &f)
{
    std::ifstream s(f);
    return s.good();
}
static bool dirExists(const std::string &d)
{
    struct stat st;
    return stat(d.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
[[noreturn]] static void fatal(const std::string &msg)
{
    // FatalError << ... << abort(FatalError)
    fprintf(stderr, "\n\n--> FOAM FATAL ERROR: \n%s\n\nFOAM aborting\n\n", msg.c_str());
    exit(1);
}

// OpenFOAM Switch parsing for bool options
static bool parseSwitch(const std::string &v, const std::string &opt)
{
    std::string s = v;
    std::transform(s.begin(), s.end(), s.begin(), ::tolower);
    if (s == "true" || s == "on" || s == "yes" || s == "y" || s == "t" || s == "1")
        return true;
    if (s == "false" || s == "off" || s == "no" || s == "n" || s == "f" || s == "none" || s == "0")
        return false;
    fatal("bad bool value '" + v + "' for option -" + opt);
}

// minimal controlDict reader: `key value;` pairs at top level
static std::map<std::string, std::string> readDict(const std::string &file)
{
    std::map<std::string, std::string> d;
    std::ifstream in(file);
    if (!in)
        return d;
    std::stringstream ss;
    ss << in.rdbuf();
    std::string s = ss.str(), clean;
    for (size_t i = 0; i < s.size(); ++i)
    { // strip comments
        if (s.compare(i, 2, "//") == 0)
        {
            while (i < s.size() && s[i] != '\n')
                ++i;
        }
        else if (s.compare(i, 2, "/*") == 0)
        {
            i = s.find("*/", i + 2);
            if (i == std::string::npos)
                break;
            ++i;
        }
        else
            clean += s[i];
    }
    int depth = 0;
    std::string stmt;
    for (char c : clean)
    {
        if (c == '{')
            ++depth;
        else if (c == '}')
        {
            --depth;
            stmt.clear();
        }
        else if (c == ';')
        {
            if (depth == 0)
            {
                std::istringstream is(stmt);
                std::string k, v;
                is >> k;
                std::getline(is, v);
                const size_t b = v.find_first_not_of(" \t\n");
                if (!k.empty() && b != std::string::npos)
                    d[k] = v.substr(b, v.find_last_not_of(" \t\n") - b + 1);
            }
            stmt.clear();
        }
        else
            stmt += c;
    }
    return d;
}

// OpenFOAM `timeFormat general; timePrecision 6;` (testcase/system/controlDict:34-36)
static std::string timeName(double t, int precision)
{
    char buf[64];
    snprintf(buf, sizeof buf, "%.*g", precision, t);
    return buf;
}

static std::vector<std::pair<double, std::string>> findTimes(const std::string &caseDir)
{
    std::vector<std::pair<double, std::string>> out;
    DIR *dp = opendir(caseDir.c_str());
    if (!dp)
        return out;
    while (dirent *e = readdir(dp))
    {
        const std::string n = e->d_name;
        char *end;
        const double v = strtod(n.c_str(), &end);
        if (end != n.c_str() && *end == '\0' && dirExists(caseDir + "/" + n))
            out.push_back({v, n});
    }
    closedir(dp);
    std::sort(out.begin(), out.end());
    return out;
}

// OpenFOAM wordRe list, e.g.  walls   '(walls "rotor.*")'   '("def.*")' : quoted entries are regular
// expressions (full match), unquoted ones literal names (getPatchIdsForOption, src/smoothMesh.C:1442-1471)
static std::vector<int32_t> selectPatches(const std::string &expr, const std::vector<std::string> &names)
{
    std::vector<int32_t> sel(names.size(), 0);
    std::string s = expr;
    for (char &c : s)
        if (c == '(' || c == ')')
            c = ' ';
    size_t i = 0;
    while (i < s.size())
    {
        while (i < s.size() && isspace((unsigned char)s[i]))
            ++i;
        if (i >= s.size())
            break;
        bool isRe = false;
        std::string tok;
        if (s[i] == '"')
        {
            isRe = true;
            const size_t e = s.find('"', i + 1);
            tok = s.substr(i + 1, (e == std::string::npos ? s.size() : e) - i - 1);
            i = (e == std::string::npos) ? s.size() : e + 1;
        }
        else
        {
            const size_t b = i;
            while (i < s.size() && !isspace((unsigned char)s[i]))
                ++i;
            tok = s.substr(b, i - b);
        }
        for (size_t k = 0; k < names.size(); ++k)
        {
            bool hit = false;
            if (isRe)
            {
                try
                {
                    hit = std::regex_match(names[k], std::regex(tok, std::regex::extended));
                }
                catch (const std::regex_error &)
                {
                    fatal("bad regular expression \"" + tok + "\" in patch list");
                }
            }
            else
                hit = names[k] == tok;
            if (hit)
                sel[k] = 1;
        }
    }
    return sel;
}

struct RunOptions
{
    smgpu_params prm;
    int centroidalIters, writeInterval;
    double deltaT, startTime;
    bool binary;
    int writePrecision, timePrecision;
    std::string caseDir, startName;
    std::string layerPatches; // -layerPatches expression, empty = none
    int logPrecision = 6;
    // boundary point smoothing (src/smoothMesh.C:2080-2171): requested when constant/geometry holds the inputs
    bool boundaryRequested = false;
    std::string smoothingPatches; // -smoothingPatches expression, empty = all patches (:1837-1840)
    double internalFraction = 0.0;
};

namespace
{
// the arrays of constant/geometry/*.obj behind an smgpu_boundary_geometry
struct BoundaryInputs
{
    struct Obj
    {
        std::vector<double> p;
        std::vector<int32_t> e, t;
    };
    Obj surf, init, target;
    bool haveTarget = false;
    smgpu_boundary_geometry geo;
    static Obj load(const std::string &f)
    {
        Obj o;
        int64_t np = 0, ne = 0, nt = 0;
        if (smmesh_read_obj(f.c_str(), &np, nullptr, &ne, nullptr, &nt, nullptr) != SMGPU_OK)
            fatal(smmesh_last_error());
        o.p.resize(3 * np);
        o.e.resize(2 * ne);
        o.t.resize(3 * nt);
        smmesh_read_obj(f.c_str(), nullptr, o.p.data(), nullptr, o.e.data(), nullptr, o.t.data());
        return o;
    }
    void read(const std::string &caseDir)
    {
        surf = load(caseDir + "/constant/geometry/targetSurfaces.obj");
        init = load(caseDir + "/constant/geometry/initEdges.obj");
        target = init;
        haveTarget = fileExists(caseDir + "/constant/geometry/targetEdges.obj");
        if (haveTarget)
            target = load(caseDir + "/constant/geometry/targetEdges.obj");
        memset(&geo, 0, sizeof geo);
        geo.n_init_points = (int64_t)init.p.size() / 3, geo.init_points = init.p.data();
        geo.n_init_edges = (int64_t)init.e.size() / 2, geo.init_edges = init.e.data();
        geo.n_target_points = (int64_t)target.p.size() / 3, geo.target_points = target.p.data();
        geo.n_target_edges = (int64_t)target.e.size() / 2, geo.target_edges = target.e.data();
        geo.n_surface_points = (int64_t)surf.p.size() / 3, geo.surface_points = surf.p.data();
        geo.n_surface_tris = (int64_t)surf.t.size() / 3, geo.surface_tris = surf.t.data();
    }
};
} // namespace

namespace
{
struct Barrier
{
    std::mutex m;
    std::condition_variable cv;
    int n, count = 0, gen = 0;
    explicit Barrier(int n_) : n(n_) {}
    void wait()
    {
        std::unique_lock<std::mutex> lk(m);
        const int g = gen;
        if (++count == n)
        {
            count = 0;
            ++gen;
            cv.notify_all();
        }
        else
            cv.wait(lk, [&] { return gen != g; });
    }
};
} // namespace

// -parallel: the decomposed case (processor<k>/, decomposePar layout) on one GPU per processor
// directory.  The reference runs one MPI rank per directory (testcase/run_parallel); here one host
// thread per directory drives its GPU and the library's NCCL layer replaces syncTools/Pstream.
static int runParallel(const RunOptions &ro)
{
    int nProcs = 0;
    while (dirExists(ro.caseDir + "/processor" + std::to_string(nProcs)))
        ++nProcs;
    if (nProcs < 2)
        fatal("-parallel: no processor directories found in " + ro.caseDir + " (decompose the case first, e.g. -decompose '(2 1 1)')");
    int32_t nDev = 0;
    smgpu_device_count(&nDev);
    if (nDev < 1)
        fatal("-parallel: no CUDA device available (there is no CPU fallback)");
    // fewer GPUs than processor directories: all processor meshes become an in-process group on one device
    // (smgpu_group_*: same rank semantics, device copies instead of NCCL)
    const bool groupMode = nDev < nProcs;
    if (groupMode)
        printf("Running %d processor meshes as an in-process group on GPU %d\n\nCreate mesh for time = %s\n\n", nProcs,
               ro.prm.device, ro.startName.c_str());
    else
        printf("Running on %d GPUs (one per processor directory)\n\nCreate mesh for time = %s\n\n", nProcs, ro.startName.c_str());

    std::vector<std::vector<int64_t>> shared(nProcs);
    std::vector<int64_t> counts(nProcs), allGids;
    uint8_t uid[128];
    Barrier bar(nProcs);
    std::mutex failMutex;
    std::string failure;
    std::atomic<bool> failed{false}, p2pFailed{false}; // what the worker threads poll; `failure` itself is only touched under the mutex
    std::vector<smgpu_handle *> handles(nProcs, nullptr);
    auto failAll = [&](const std::string &msg) {
        std::lock_guard<std::mutex> lk(failMutex);
        if (failure.empty())
            failure = msg;
        if (!failed.exchange(true) && !groupMode)
            for (smgpu_handle *q : handles) // peers may sit in an NCCL kernel waiting for the rank that failed
                if (q)
                    smgpu_comm_abort(q);
    };
    BoundaryInputs boundary;
    if (ro.boundaryRequested)
    {
        boundary.read(ro.caseDir);
        printf("Enabled boundary point smoothing\n\n");
    }
    std::vector<std::vector<int32_t>> smoothSel(nProcs);
    smgpu_group *group = nullptr;
    int sharedDone = 0;
    bool sharedOk = true;
    std::vector<int64_t> sharedFrozen(std::max(ro.centroidalIters, 1));
    std::vector<double> sharedResidual(std::max(ro.centroidalIters, 1));
    int64_t totalPoints = 0, totalInternal = 0;
    std::vector<std::thread> threads;
    for (int k = 0; k < nProcs; ++k)
        threads.emplace_back([&, k] {
            const std::string procDir = ro.caseDir + "/processor" + std::to_string(k);
            smmesh *mesh = smmesh_read_processor(ro.caseDir.c_str(), k);
            smgpu_handle *h = nullptr;
            bool ok = mesh != nullptr;
            if (!ok)
                failAll(std::string("cannot read ") + procDir + ": " + smmesh_last_error());
            if (ok && ro.startName != "constant" && fileExists(procDir + "/" + ro.startName + "/polyMesh/points"))
                if (smmesh_read_points(mesh, (procDir + "/" + ro.startName + "/polyMesh/points").c_str()) != SMGPU_OK)
                {
                    ok = false;
                    failAll(smmesh_last_error());
                }
            int64_t nPoints = 0;
            std::vector<int32_t> pStart, pSize, pKind, layerSel;
            if (ok)
            {
                nPoints = smmesh_size(mesh, 0);
                const int nPatches = (int)smmesh_size(mesh, 5);
                pStart.resize(nPatches), pSize.resize(nPatches), pKind.resize(nPatches);
                layerSel.assign(nPatches, 0);
                if (!ro.layerPatches.empty())
                {
                    // patchSet on this processor's boundary file, like every MPI rank of the reference (:1826-1833)
                    std::vector<std::string> names;
                    for (int i = 0; i < nPatches; ++i)
                        names.push_back(smmesh_patch_name(mesh, i));
                    layerSel = selectPatches(ro.layerPatches, names);
                }
                if (ro.boundaryRequested)
                {
                    std::vector<std::string> names;
                    for (int i = 0; i < nPatches; ++i)
                        names.push_back(smmesh_patch_name(mesh, i));
                    smoothSel[k] = ro.smoothingPatches.empty() ? std::vector<int32_t>(nPatches, 1) : selectPatches(ro.smoothingPatches, names);
                }
                smmesh_patches(mesh, pStart.data(), pSize.data(), pKind.data());
                smgpu_mesh_desc md;
                memset(&md, 0, sizeof md);
                md.n_points = nPoints;
                md.n_cells = smmesh_size(mesh, 1);
                md.n_faces = smmesh_size(mesh, 2);
                md.n_internal_faces = smmesh_size(mesh, 3);
                md.points = smmesh_points(mesh);
                md.face_offsets = smmesh_face_offsets(mesh);
                md.face_verts = smmesh_face_verts(mesh);
                md.owner = smmesh_owner(mesh);
                md.neighbour = smmesh_neighbour(mesh);
                md.n_patches = nPatches;
                md.patch_start = pStart.data();
                md.patch_size = pSize.data();
                md.patch_kind = pKind.data();
                md.point_global_id = smmesh_point_global_id(mesh);
                md.patch_layer = layerSel.data();
                smgpu_params prm = ro.prm;
                prm.device = groupMode ? ro.prm.device : k;
                if (smgpu_create(&md, &prm, &h) != SMGPU_OK)
                {
                    ok = false;
                    failAll(smgpu_last_error());
                }
                handles[k] = h;
            }
            if (ok)
            {
                int64_t n = 0;
                if (smgpu_comm_local_shared(h, &n, nullptr) != SMGPU_OK)
                    failAll(smgpu_last_error());
                shared[k].resize(n);
                smgpu_comm_local_shared(h, &n, shared[k].data());
            }
            bar.wait();
            if (k == 0 && !failed && groupMode)
            {
                if (smgpu_group_create(handles.data(), nProcs, &group) != SMGPU_OK)
                    failAll(smgpu_last_error());
                else if (ro.boundaryRequested)
                {
                    std::vector<const int32_t *> sel;
                    for (auto &v : smoothSel)
                        sel.push_back(v.data());
                    if (smgpu_group_enable_boundary_smoothing(group, &boundary.geo, sel.data(), ro.internalFraction) != SMGPU_OK)
                        failAll(smgpu_last_error());
                }
            }
            else if (k == 0 && !failed)
            {
                for (int r = 0; r < nProcs; ++r)
                {
                    counts[r] = (int64_t)shared[r].size();
                    allGids.insert(allGids.end(), shared[r].begin(), shared[r].end());
                }
                if (allGids.empty())
                    allGids.push_back(0);
                if (smgpu_comm_unique_id(uid) != SMGPU_OK)
                    failAll(smgpu_last_error());
            }
            bar.wait();
            if (failed)
                return;
            if (!groupMode)
            {
                // local half first, so that a rank whose plan fails does not leave the others in ncclCommInitRank
                if (smgpu_comm_prepare(h, k, nProcs, counts.data(), allGids.data()) != SMGPU_OK)
                    failAll(smgpu_last_error());
                bar.wait();
                if (failed)
                    return;
                if (smgpu_comm_init(h, k, nProcs, uid, counts.data(), allGids.data()) != SMGPU_OK)
                    failAll(smgpu_last_error());
                bar.wait();
                if (failed)
                    return;
                // boundary point smoothing: collective set-up on every rank (:2080-2250 under -parallel)
                if (ro.boundaryRequested &&
                    smgpu_enable_boundary_smoothing(h, &boundary.geo, smoothSel[k].data(), ro.internalFraction) != SMGPU_OK)
                    failAll(smgpu_last_error());
                bar.wait();
                if (failed)
                    return;
                // peer-memory exchange between the GPUs of this process (direct peer access); all ranks or none
                if (smgpu_comm_p2p_connect(h, nullptr, handles.data()) != SMGPU_OK)
                    p2pFailed = true;
                bar.wait();
                if (p2pFailed)
                    smgpu_comm_p2p_disable(h);
            }
            bar.wait();
            if (failed)
                return;
            double mn, mx;
            int64_t nInternal, nEdges;
            smgpu_mesh_stats(h, &mn, &mx, &nInternal, &nEdges);
            {
                std::lock_guard<std::mutex> lk(failMutex);
                totalPoints += nPoints;
                totalInternal += nInternal;
            }
            bar.wait();
            smgpu_params prm;
            smgpu_get_params(h, &prm);
            if (k == 0)
            {
                printf("Applying following parameter values in smoothing:\n    centroidalIters        %d\n    relTol"
                       "                 %g\n    minEdgeLength          %g\n    maxStepLength          %g\n    "
                       "relStepFrac            %g\n    totalMinFreeze         %d\n\n",
                       ro.centroidalIters, prm.rel_tol, prm.min_edge_length, prm.max_step_length, prm.rel_step_frac,
                       prm.total_min_freeze);
                printf("Mesh includes a total of %lld points:\n  - %lld internal (non-boundary) points\n  - %lld boundary "
                       "points\nMesh minimum edge length = %g\nMesh maximum edge length = %g\n\n",
                       (long long)totalPoints, (long long)totalInternal, (long long)(totalPoints - totalInternal), mn, mx);
            }
            std::vector<double> pts(3 * nPoints);
            std::vector<int64_t> nFrozen(std::max(ro.centroidalIters, 1));
            std::vector<double> residual(std::max(ro.centroidalIters, 1));
            int i = 0;
            const int logPrecision = ro.logPrecision;
            bool stop = ro.centroidalIters <= 0;
            while (!stop && (groupMode || !failed))
            {
                int chunk = ro.centroidalIters - i;
                if (ro.writeInterval > 0)
                {
                    int toWrite = ro.writeInterval - (i % ro.writeInterval);
                    if (i + toWrite - 1 == 0)
                        toWrite += ro.writeInterval; // the `i > 0` quirk at :2416
                    chunk = std::min(chunk, toWrite);
                }
                int done = 0;
                if (groupMode)
                { // one thread drives the whole group, the others take its log (the chunk is the same on every thread)
                    if (k == 0)
                    {
                        sharedOk = smgpu_group_iterate(group, chunk, sharedFrozen.data(), sharedResidual.data(), &sharedDone) == SMGPU_OK;
                        if (!sharedOk)
                            failAll(smgpu_last_error());
                    }
                    bar.wait();
                    const bool okNow = sharedOk;
                    done = sharedDone;
                    std::copy(sharedFrozen.begin(), sharedFrozen.begin() + done, nFrozen.begin());
                    std::copy(sharedResidual.begin(), sharedResidual.begin() + done, residual.begin());
                    bar.wait();
                    if (!okNow)
                        break;
                }
                else if (smgpu_iterate(h, chunk, nFrozen.data(), residual.data(), &done) != SMGPU_OK)
                {
                    failAll(smgpu_last_error());
                    break;
                }
                if (k == 0)
                    for (int q = 0; q < done; ++q)
                        printf("Smoothing iteration=%d nFrozenPoints=%lld residual=%.*g\n", i + q + 1, (long long)nFrozen[q],
                               logPrecision, residual[q]);
                i += done;
                if (done > 0 && residual[done - 1] < prm.rel_tol)
                {
                    if (k == 0)
                        printf("Residual reached relTol, stopping.\n");
                    stop = true;
                }
                if (i >= ro.centroidalIters)
                {
                    if (k == 0)
                        printf("Maximum centroidalIters reached, stopping.\n");
                    stop = true;
                }
                const int last = i - 1;
                if (stop || (ro.writeInterval > 0 && ((last + 1) % ro.writeInterval) == 0 && last > 0))
                {
                    const std::string tn = timeName(ro.startTime + i * ro.deltaT, ro.timePrecision);
                    if (k == 0)
                        printf("Writing new mesh to time %s\n\n", tn.c_str());
                    if (smgpu_get_points(h, pts.data()) != SMGPU_OK ||
                        smmesh_write_points(pts.data(), nPoints, (procDir + "/" + tn + "/polyMesh").c_str(), ro.binary,
                                            std::max(10, ro.writePrecision), (tn + "/polyMesh").c_str()) != SMGPU_OK)
                        failAll("cannot write points of processor " + std::to_string(k));
                }
            }
            bar.wait(); // nobody tears its communicator down while others still iterate
            if (groupMode)
            {
                if (k == 0)
                    smgpu_group_destroy(group);
                bar.wait();
            }
            smgpu_destroy(h);
            smmesh_free(mesh);
        });
    for (auto &t : threads)
        t.join();
    if (!failure.empty())
        fatal(failure);
    printf("\nEnd\n");
    return 0;
}

int main(int argc, char **argv)
{
    const time_t wallStart = time(nullptr);
    // ---- option table: name -> has value (src/smoothMesh.C:1642-1784 + OpenFOAM standard options)
    const char *valued[] = {"case",
                            "time",
                            "centroidalIters",
                            "maxStepLength",
                            "relStepFrac",
                            "edgeAngleConstraint",
                            "faceAngleConstraint",
                            "minEdgeLength",
                            "totalMinFreeze",
                            "minAngle",
                            "maxAngle",
                            "layerMaxBlendingFraction",
                            "layerEdgeLength",
                            "layerExpansionRatio",
                            "minLayers",
                            "maxLayers",
                            "layerPatches",
                            "smoothingPatches",
                            "internalSmoothingBlendingFraction",
                            "relTol",
                            "writeInterval",
                            "device",
                            "geometryVariant",
                            "decompose"};
    std::map<std::string, std::string> opt;
    bool parallel = false;
    for (int i = 1; i < argc; ++i)
    {
        std::string a = argv[i];
        if (a.size() < 2 || a[0] != '-')
            fatal("Wrong number of arguments, expected 0 found 1\nInvalid argument: " + a);
        a = a.substr(1);
        if (a == "parallel")
        {
            parallel = true;
            continue;
        }
        if (a == "help")
        {
            printf("Usage: smoothMesh [OPTIONS]\nMove internal mesh points to increase mesh quality\n"
                   "options: -case <dir> -time <time> -centroidalIters <label> -relTol <double> -minEdgeLength <double>\n"
                   "  -maxStepLength <double> -relStepFrac <double> -totalMinFreeze <bool> -edgeAngleConstraint <bool>\n"
                   "  -faceAngleConstraint <bool> -minAngle <double> -maxAngle <double> -writeInterval <label>\n"
                   "  -layerPatches <wordRe> -smoothingPatches <wordRe> (only values that keep those features off)\n"
                   "  -device <int> -geometryVariant com|org\n");
            return 0;
        }
        bool known = false;
        for (const char *v : valued)
            known = known || a == v;
        if (!known)
            fatal("Invalid option: -" + a);
        if (i + 1 >= argc)
            fatal("Option -" + a + " requires a value");
        opt[a] = argv[++i];
    }
    auto has = [&](const char *k) { return opt.count(k) > 0; };
    auto num = [&](const char *k, double dflt) { return has(k) ? atof(opt[k].c_str()) : dflt; };

    printf("smoothMesh (smoothmesh_b200: %s)\n\n", smgpu_version());
    const std::string caseDir = has("case") ? opt["case"] : ".";
    std::map<std::string, std::string> control = readDict(caseDir + "/system/controlDict");
    const double deltaT = control.count("deltaT") ? atof(control["deltaT"].c_str()) : 1.0;
    if (deltaT < 1e-300) // src/smoothMesh.C:1806-1812
        fatal("Time step (deltaT) value " + std::to_string(deltaT) + " specified in controlDict is too small");
    const bool binary = control.count("writeFormat") && control["writeFormat"] == "binary";
    const int writePrecision = control.count("writePrecision") ? atoi(control["writePrecision"].c_str()) : 6;
    const int timePrecision = control.count("timePrecision") ? atoi(control["timePrecision"].c_str()) : 6;
    // [OF-recalled] Time::readDict: a writePrecision entry also sets the precision of the log streams
    const int lp = control.count("writePrecision") ? writePrecision : 6;

    // ---- start time (createTime.H + :1792-1803) and mesh instance lookup (createMesh.H)
    // -parallel looks the time directories up under processor0/ (the case root holds none after decomposePar)
    std::vector<std::pair<double, std::string>> times = findTimes(parallel ? caseDir + "/processor0" : caseDir);
    double startTime = 0;
    std::string startName = "0";
    if (has("time"))
    {
        if (opt["time"] == "constant")
        {
            startTime = 0;
            startName = "constant";
        }
        else
        {
            startTime = atof(opt["time"].c_str());
            startName = timeName(startTime, timePrecision);
        }
    }
    else
    {
        const std::string startFrom = control.count("startFrom") ? control["startFrom"] : "latestTime";
        if (startFrom == "latestTime" && !times.empty())
        {
            startTime = times.back().first;
            startName = times.back().second;
        }
        else if (startFrom == "firstTime" && !times.empty())
        {
            startTime = times.front().first;
            startName = times.front().second;
        }
        else
        {
            startTime = control.count("startTime") ? atof(control["startTime"].c_str()) : 0.0;
            startName = timeName(startTime, timePrecision);
        }
    }
    // ---- options -> library parameters (:1863-1931)
    const double layerMaxBlendingFraction = num("layerMaxBlendingFraction", 0.3);
    smgpu_params prm;
    smgpu_default_params(&prm);
    prm.min_edge_length = num("minEdgeLength", -1.0);
    prm.max_step_length = num("maxStepLength", -1.0);
    prm.rel_step_frac = num("relStepFrac", 0.5);
    prm.total_min_freeze = has("totalMinFreeze") ? parseSwitch(opt["totalMinFreeze"], "totalMinFreeze") : 0;
    prm.min_angle_deg = num("minAngle", 35.0);
    prm.max_angle_deg = num("maxAngle", 160.0);
    prm.edge_angle_constraint = has("edgeAngleConstraint") ? parseSwitch(opt["edgeAngleConstraint"], "edgeAngleConstraint") : 1;
    prm.face_angle_constraint = has("faceAngleConstraint") ? parseSwitch(opt["faceAngleConstraint"], "faceAngleConstraint") : 1;
    prm.rel_tol = num("relTol", 0.02);
    prm.device = (int)num("device", 0);
    prm.geometry_variant = has("geometryVariant") && opt["geometryVariant"] == "org" ? 1 : 0;
    prm.layer_max_blending_fraction = layerMaxBlendingFraction;
    prm.layer_edge_length = num("layerEdgeLength", -1.0);
    prm.layer_expansion_ratio = num("layerExpansionRatio", 1.3);
    prm.min_layers = (int)num("minLayers", 1);
    prm.max_layers = (int)num("maxLayers", 4);
    const int centroidalIters = (int)num("centroidalIters", 1000);
    const int writeInterval = (int)num("writeInterval", centroidalIters);
    auto patchSetEmpty = [&](const std::string &expr) {
        std::string s = expr;
        s.erase(std::remove_if(s.begin(), s.end(), [](char c) { return isspace((unsigned char)c) || c == '"'; }), s.end());
        return s == "()" || s == "none" || s == "(none)" || s.empty();
    };
    const bool smoothingPatchesEmpty = has("smoothingPatches") && patchSetEmpty(opt["smoothingPatches"]);
    const bool surfaces = fileExists(caseDir + "/constant/geometry/targetSurfaces.obj");
    const bool initEdges = fileExists(caseDir + "/constant/geometry/initEdges.obj");
    // :2080-2086 (the isCornerPoint / isFeatureEdgePoint label lists of earlier runs are not read)
    // :2039-2078: the classification lists an earlier run left in the start time directory
    auto readList = [&](const std::string &name) {
        std::vector<int32_t> v;
        const std::string f = caseDir + "/" + startName + "/" + name;
        if (!parallel && fileExists(f))
        {
            const int64_t n = smmesh_read_label_list(f.c_str(), nullptr, 0);
            if (n > 0)
            {
                v.resize(n);
                smmesh_read_label_list(f.c_str(), v.data(), n);
            }
        }
        return v;
    };
    const std::vector<int32_t> cornerIO = readList("isCornerPoint"), featureIO = readList("isFeatureEdgePoint");
    const bool labelIOListsHaveData = std::find(cornerIO.begin(), cornerIO.end(), 1) != cornerIO.end() ||
                                      std::find(featureIO.begin(), featureIO.end(), 1) != featureIO.end();
    const bool boundaryRequested = !has("decompose") && surfaces && (initEdges || labelIOListsHaveData) && !smoothingPatchesEmpty;

    if (parallel)
    {
        if (has("decompose"))
            fatal("-decompose and -parallel are separate steps");
        RunOptions ro;
        if (has("layerPatches") && !patchSetEmpty(opt["layerPatches"]))
            ro.layerPatches = opt["layerPatches"];
        ro.prm = prm;
        ro.logPrecision = lp;
        ro.centroidalIters = centroidalIters;
        ro.writeInterval = writeInterval;
        ro.deltaT = deltaT;
        ro.startTime = startTime;
        ro.binary = binary;
        ro.writePrecision = writePrecision;
        ro.timePrecision = timePrecision;
        ro.caseDir = caseDir;
        ro.startName = startName;
        ro.boundaryRequested = boundaryRequested;
        if (has("smoothingPatches"))
            ro.smoothingPatches = opt["smoothingPatches"];
        ro.internalFraction = num("internalSmoothingBlendingFraction", 0.0);
        printf("Create time\n\n");
        return runParallel(ro);
    }

    // newest time <= start time that has polyMesh/points; topology from the newest with polyMesh/faces; else constant
    std::string pointsDir = caseDir + "/constant/polyMesh", topoDir = caseDir + "/constant/polyMesh";
    if (startName != "constant")
        for (auto &t : times)
            if (t.first <= startTime * (1 + 1e-12) + 1e-300)
            {
                if (fileExists(caseDir + "/" + t.second + "/polyMesh/points"))
                    pointsDir = caseDir + "/" + t.second + "/polyMesh";
                if (fileExists(caseDir + "/" + t.second + "/polyMesh/faces"))
                    topoDir = caseDir + "/" + t.second + "/polyMesh";
            }
    printf("Create time\n\nCreate mesh for time = %s\n\n", startName.c_str());
    smmesh *mesh = smmesh_read(topoDir.c_str());
    if (!mesh)
        fatal(std::string("cannot read polyMesh from ") + topoDir + ": " + smmesh_last_error());
    if (pointsDir != topoDir)
    {
        // points from a later time directory (mesh.write() only writes points)
        if (smmesh_read_points(mesh, (pointsDir + "/points").c_str()) != SMGPU_OK)
            fatal(std::string("cannot read ") + pointsDir + "/points: " + smmesh_last_error());
    }
    if (has("decompose"))
    {
        // utility mode standing in for decomposePar (testcase/run_parallel:19): simple (nx ny nz) bricks
        int d[3] = {0, 0, 0};
        std::string v = opt["decompose"];
        for (char &c : v)
            if (c == '(' || c == ')' || c == ',')
                c = ' ';
        if (sscanf(v.c_str(), "%d %d %d", &d[0], &d[1], &d[2]) != 3 || d[0] < 1 || d[1] < 1 || d[2] < 1)
            fatal("-decompose expects '(nx ny nz)'");
        const int nParts = d[0] * d[1] * d[2];
        std::vector<smmesh *> parts(nParts, nullptr);
        if (smmesh_decompose(mesh, 0, d[0], d[1], d[2], parts.data()) != SMGPU_OK ||
            smmesh_write_decomposed(parts.data(), nParts, caseDir.c_str(), binary) != SMGPU_OK)
            fatal(smmesh_last_error());
        for (smmesh *q : parts)
            smmesh_free(q);
        printf("Decomposed mesh of time %s into %d processor directories under %s\n\nEnd\n", startName.c_str(),
               d[0] * d[1] * d[2], caseDir.c_str());
        smmesh_free(mesh);
        return 0;
    }
    const int64_t nPoints = smmesh_size(mesh, 0);
    const int nPatches = (int)smmesh_size(mesh, 5);
    std::vector<int32_t> pStart(nPatches), pSize(nPatches), pKind(nPatches);
    smmesh_patches(mesh, pStart.data(), pSize.data(), pKind.data());

    // ---- features outside the hot path (:1823-1852, :2024-2098)
    std::vector<std::string> patchNames;
    for (int i = 0; i < nPatches; ++i)
        patchNames.push_back(smmesh_patch_name(mesh, i));
    std::vector<int32_t> layerSel(nPatches, 0);
    bool anyLayerPatch = false;
    if (has("layerPatches"))
    {
        layerSel = selectPatches(opt["layerPatches"], patchNames);
        for (int32_t f : layerSel)
            anyLayerPatch = anyLayerPatch || f;
    }
    if (anyLayerPatch)
        printf("Patches for boundary layer treatment: %s\n", opt["layerPatches"].c_str());
    else
        printf("Patches for boundary layer treatment: none\n");
    const bool doLayerTreatment = anyLayerPatch && layerMaxBlendingFraction > 1e-15; // :2025
    {
        // :1837-1854: the option defaults to '(".*")'; "none" is printed when no patch matches
        bool anySmoothing = !has("smoothingPatches");
        if (has("smoothingPatches"))
            for (int32_t f : selectPatches(opt["smoothingPatches"], patchNames))
                anySmoothing = anySmoothing || f;
        printf("Patches for boundary point smoothing: %s\n",
               !anySmoothing ? "none" : (has("smoothingPatches") ? opt["smoothingPatches"].c_str() : "(\".*\")"));
    }

    // ---- GPU handle
    smgpu_mesh_desc md;
    memset(&md, 0, sizeof md);
    md.n_points = nPoints;
    md.n_cells = smmesh_size(mesh, 1);
    md.n_faces = smmesh_size(mesh, 2);
    md.n_internal_faces = smmesh_size(mesh, 3);
    md.points = smmesh_points(mesh);
    md.face_offsets = smmesh_face_offsets(mesh);
    md.face_verts = smmesh_face_verts(mesh);
    md.owner = smmesh_owner(mesh);
    md.neighbour = smmesh_neighbour(mesh);
    md.n_patches = nPatches;
    md.patch_start = pStart.data();
    md.patch_size = pSize.data();
    md.patch_kind = pKind.data();
    md.patch_layer = layerSel.data();

    smgpu_handle *h = nullptr;
    if (smgpu_create(&md, &prm, &h) != SMGPU_OK)
        fatal(smgpu_last_error());
    smgpu_get_params(h, &prm);
    double meshMinEdge, meshMaxEdge;
    int64_t nInternal, nEdges;
    smgpu_mesh_stats(h, &meshMinEdge, &meshMaxEdge, &nInternal, &nEdges);
    if (prm.max_step_length > 0.5 * prm.min_edge_length)
        printf("WARNING: The maximum allowed step length is more than half of the minimum edge length! This may cause "
               "unstability in smoothing.\n\n");

    // parameter echo, :1933-1975
    printf("Applying following parameter values in smoothing:\n");
    printf("    centroidalIters        %d\n", centroidalIters);
    printf("    relTol                 %.*g\n", lp, prm.rel_tol);
    printf("    minEdgeLength          %.*g\n", lp, prm.min_edge_length);
    printf("    maxStepLength          %.*g\n", lp, prm.max_step_length);
    printf("    relStepFrac            %.*g\n", lp, prm.rel_step_frac);
    printf("    totalMinFreeze         %d\n", prm.total_min_freeze);
    if (prm.edge_angle_constraint)
        printf("    edgeAngleConstraint    true\n    minAngle               %.*g\n", lp, prm.min_angle_deg);
    else
        printf("    edgeAngleConstraint    false (edge min angle quality constraint is NOT applied)\n");
    if (prm.face_angle_constraint)
        printf("    faceAngleConstraint    true\n    minAngle               %.*g\n    maxAngle               %.*g\n", lp,
               prm.min_angle_deg, lp, prm.max_angle_deg);
    else
        printf("    faceAngleConstraint    false (face angle quality constraints are NOT applied)\n");
    if (layerMaxBlendingFraction > 1e-15)
        printf("    layerMaxBlendingFraction %.*g\n    layerEdgeLength          %.*g\n    layerExpansionRatio      %.*g\n"
               "    minLayers                %d\n    maxLayers                %d\n\n",
               lp, prm.layer_max_blending_fraction, lp,
               prm.layer_edge_length < 0 ? prm.min_edge_length : prm.layer_edge_length, lp, prm.layer_expansion_ratio,
               prm.min_layers, prm.max_layers);
    else
        printf("    layerMaxBlendingFraction 0 (boundary layer treatment is NOT applied)\n\n");
    // the remaining set-up messages in the reference's order (:2010-2012, :2027-2098, :2181-2187 and
    // src/boundaryPointSmoothing.C:432-438), so that logs of the two tools can be compared line by line
    printf("Starting to build pointNeighPoints (this may take some time)\nDone building pointNeighPoints\n\n");
    if (doLayerTreatment)
        printf("Enabled boundary layer treatment\n\n");
    else
        printf("Boundary layer treatment is disabled. Either no layerPatches were specified or "
               "boundaryMaxBlendingFraction is zero\n\n");
    if (labelIOListsHaveData)
        printf("Found corners and feature edges in isCornerPoint and isFeatureEdgePoint files\n\n");
    else
        printf("Did not find corners and feature edges in isCornerPoint and isFeatureEdgePoint files\n\n");
    // boundary point smoothing, :2080-2171: inputs from constant/geometry, set-up inside the library
    std::vector<int32_t> smoothSel(nPatches, 0);
    if (boundaryRequested)
        smoothSel = has("smoothingPatches") ? selectPatches(opt["smoothingPatches"], patchNames)
                                            : std::vector<int32_t>(nPatches, 1); // default '(".*")', :1837-1840
    bool doBoundarySmoothing = false;
    for (int32_t f : smoothSel)
        doBoundarySmoothing = doBoundarySmoothing || f;
    if (doBoundarySmoothing)
    {
        printf("Enabled boundary point smoothing\n\n");
        struct Obj
        {
            std::vector<double> p;
            std::vector<int32_t> e, t;
        };
        auto load = [&](const std::string &rel) {
            Obj o;
            int64_t np = 0, ne = 0, nt = 0;
            const std::string f = caseDir + "/" + rel;
            if (smmesh_read_obj(f.c_str(), &np, nullptr, &ne, nullptr, &nt, nullptr) != SMGPU_OK)
                fatal(smmesh_last_error());
            o.p.resize(3 * np);
            o.e.resize(2 * ne);
            o.t.resize(3 * nt);
            smmesh_read_obj(f.c_str(), nullptr, o.p.data(), nullptr, o.e.data(), nullptr, o.t.data());
            return o;
        };
        // (the statistics lines are this tool's; OpenFOAM's triSurface / edgeMesh writeStats print more)
        const Obj surf = load("constant/geometry/targetSurfaces.obj");
        printf("Target surfaces file constant/geometry/targetSurfaces.obj stats:\nTriangles    : %lld\nVertices     : %lld\n\n",
               (long long)(surf.t.size() / 3), (long long)(surf.p.size() / 3));
        const Obj ie = load("constant/geometry/initEdges.obj");
        printf("Initial feature edges file constant/geometry/initEdges.obj stats:\npoints      : %lld\nedges       : %lld\n\n",
               (long long)(ie.p.size() / 3), (long long)(ie.e.size() / 2));
        Obj te = ie;
        if (fileExists(caseDir + "/constant/geometry/targetEdges.obj"))
        {
            te = load("constant/geometry/targetEdges.obj");
            printf("Target feature edges file constant/geometry/targetEdges.obj stats:\npoints      : %lld\nedges       : %lld\n",
                   (long long)(te.p.size() / 3), (long long)(te.e.size() / 2));
        }
        else
            printf("WARNING: Initial feature edges will be used also as target edges, because\ndid not find file "
                   "constant/geometry/targetEdges.obj.\n\n");
        printf("Checking initial edge mesh sanity\nChecking target edge mesh sanity\n");
        smgpu_boundary_geometry geo;
        memset(&geo, 0, sizeof geo);
        if (labelIOListsHaveData && (int64_t)cornerIO.size() == nPoints && (int64_t)featureIO.size() == nPoints)
        {
            geo.is_corner_point = cornerIO.data();
            geo.is_feature_edge_point = featureIO.data();
        }
        geo.n_init_points = (int64_t)ie.p.size() / 3, geo.init_points = ie.p.data();
        geo.n_init_edges = (int64_t)ie.e.size() / 2, geo.init_edges = ie.e.data();
        geo.n_target_points = (int64_t)te.p.size() / 3, geo.target_points = te.p.data();
        geo.n_target_edges = (int64_t)te.e.size() / 2, geo.target_edges = te.e.data();
        geo.n_surface_points = (int64_t)surf.p.size() / 3, geo.surface_points = surf.p.data();
        geo.n_surface_tris = (int64_t)surf.t.size() / 3, geo.surface_tris = surf.t.data();
        if (smgpu_enable_boundary_smoothing(h, &geo, smoothSel.data(), num("internalSmoothingBlendingFraction", 0.0)) != SMGPU_OK)
            fatal(smgpu_last_error());
        int64_t counts[4];
        smgpu_boundary_counts(h, counts);
        printf("\nStarting to build targetEdgeStrings\nDetected number of target edge mesh strings: %lld\n\n", (long long)counts[3]);
    }
    else
        printf("Boundary point smoothing is disabled. Missing smoothingPatches, or one or both of files:\n"
               "constant/geometry/targetSurfaces.obj\nconstant/geometry/initEdges.obj\n\n");
    if (doLayerTreatment && !doBoundarySmoothing)
        printf("WARNING: Boundary layer treatment will be done without boundary point smoothing. This can result in "
               "distorted boundary cells.\n\n");
    const double layerEdgeLengthEcho = prm.layer_edge_length < 0 ? prm.min_edge_length : prm.layer_edge_length;
    printf("Mesh includes a total of %lld points:\n  - %lld internal (non-boundary) points\n  - %lld boundary points\n"
           "Mesh minimum edge length = %.*g\nMesh maximum edge length = %.*g\nDistance tolerance = %.*g\n\n",
           (long long)nPoints, (long long)nInternal, (long long)(nPoints - nInternal), lp, meshMinEdge, lp, meshMaxEdge, lp,
           1e-4 * std::min(meshMinEdge, layerEdgeLengthEcho)); // REL_TOL * min(...), :1921
    {
        // classifyBoundaryPoints' summary (src/boundaryPointSmoothing.C:301-438) for a run without boundary
        // point smoothing: every boundary point is classified once, by the first patch that contains it
        const int32_t *fo = smmesh_face_offsets(mesh), *fv = smmesh_face_verts(mesh);
        std::vector<uint8_t> visited(nPoints, 0);
        long long nLayerSurface = 0, nSmoothingSurface = 0, nFrozenSurface = 0;
        for (int pi = 0; pi < nPatches; ++pi)
            for (int32_t f = pStart[pi]; f < pStart[pi] + pSize[pi]; ++f)
                for (int32_t k = fo[f]; k < fo[f + 1]; ++k)
                    if (!visited[fv[k]])
                    {
                        visited[fv[k]] = 1;
                        nLayerSurface += layerSel[pi] ? 1 : 0;
                        if (doBoundarySmoothing && smoothSel[pi])
                            ++nSmoothingSurface;
                        else
                            ++nFrozenSurface;
                    }
        int64_t counts[4] = {0, 0, 0, 0};
        if (doBoundarySmoothing)
            smgpu_boundary_counts(h, counts);
        printf("Boundary point classification summary:\n- Detected number of corner points: %lld\n- Detected number of "
               "feature edge points: %lld\n- Detected number of layer surface points: %lld\n- Detected number of "
               "smoothing surface points: %lld\n- Detected number of frozen surface points: %lld\n\n",
               (long long)counts[0], (long long)counts[1], nLayerSurface, nSmoothingSurface, nFrozenSurface);
    }

    // the two label lists are AUTO_WRITE objects of the mesh: the reference writes them with every mesh.write(),
    // all zeros when boundary point smoothing is off (:2039-2065)
    std::vector<int32_t> isCornerOut(nPoints, 0), isFeatureOut(nPoints, 0);
    if (doBoundarySmoothing)
        smgpu_get_boundary_classes(h, isCornerOut.data(), isFeatureOut.data());

    // ---- iteration loop, :2257-2437.  The library stops on relTol by itself; the loop here is
    // cut at write intervals so intermediate meshes can be written (:2416).
    std::vector<double> pts(3 * nPoints);
    std::vector<int64_t> nFrozen(std::max(centroidalIters, 1));
    std::vector<double> residual(std::max(centroidalIters, 1));
    int i = 0;
    const int logPrecision = lp;
    double totalMs = 0;
    bool stop = centroidalIters <= 0;
    while (!stop)
    {
        int chunk = centroidalIters - i;
        if (writeInterval > 0)
        {
            int toWrite = writeInterval - (i % writeInterval); // iterations until (i+1) % writeInterval == 0
            if (i + toWrite - 1 == 0)
                toWrite += writeInterval; // the `i > 0` quirk at :2416
            chunk = std::min(chunk, toWrite);
        }
        int done = 0;
        if (smgpu_iterate(h, chunk, nFrozen.data(), residual.data(), &done) != SMGPU_OK)
            fatal(smgpu_last_error());
        double ms;
        smgpu_last_timing(h, &ms, nullptr);
        totalMs += ms;
        for (int k = 0; k < done; ++k)
            printf("Smoothing iteration=%d nFrozenPoints=%lld residual=%.*g\n", i + k + 1, (long long)nFrozen[k],
                   logPrecision, residual[k]);
        i += done;
        const bool reachedTol = done > 0 && residual[done - 1] < prm.rel_tol;
        if (reachedTol)
        {
            printf("Residual reached relTol, stopping.\n");
            stop = true;
        }
        if (i >= centroidalIters)
        {
            printf("Maximum centroidalIters reached, stopping.\n");
            stop = true;
        }
        const int last = i - 1; // reference loop index of the iteration just finished
        if (stop || (writeInterval > 0 && ((last + 1) % writeInterval) == 0 && last > 0))
        {
            const std::string tn = timeName(startTime + i * deltaT, timePrecision);
            // :2425 raises IOstream::defaultPrecision for the points file; the log stream keeps its precision
            printf("Writing new mesh to time %s\n\n", tn.c_str());
            if (smgpu_get_points(h, pts.data()) != SMGPU_OK)
                fatal(smgpu_last_error());
            if (smmesh_write_points(pts.data(), nPoints, (caseDir + "/" + tn + "/polyMesh").c_str(), binary,
                                    std::max(10, writePrecision), (tn + "/polyMesh").c_str()) != SMGPU_OK)
                fatal(smmesh_last_error());
            if (smmesh_write_label_list((caseDir + "/" + tn + "/isCornerPoint").c_str(), "isCornerPoint", tn.c_str(),
                                        isCornerOut.data(), nPoints, binary) != SMGPU_OK ||
                smmesh_write_label_list((caseDir + "/" + tn + "/isFeatureEdgePoint").c_str(), "isFeatureEdgePoint", tn.c_str(),
                                        isFeatureOut.data(), nPoints, binary) != SMGPU_OK)
                fatal(smmesh_last_error());
        }
    }
    printf("GPU iteration time = %.3f ms (%d iterations, %.4g point-updates/s)\n", totalMs, i,
           totalMs > 0 ? 1e3 * double(nPoints) * i / totalMs : 0.0);
    printf("ClockTime = %lld s.\n\nEnd\n", (long long)(time(nullptr) - wallStart));
    smgpu_destroy(h);
    smmesh_free(mesh);
    return 0;
}
