// mlsgpu — command-line driver with the reference's option surface (src/mlsgpu_core.cpp:73-175,
// mlsgpu.cpp:83-263) on top of libmlsb200:
//
//   PLY files -> conditioned splats -> bounding grid -> bins -> worker group (octree, MLS,
//   marching on the GPU) -> host mesher (seam weld, prune) -> PLY
//
// Option names, defaults and validation follow the reference; options that configure parts of the
// reference that do not exist here (its reader / writer classes, reorder and gather buffers, MPI) are
// accepted and ignored so that existing command lines keep working.  Inputs larger than
// --mem-load-splats are streamed (stream.hpp); with --tmp-dir the mesher is out of core as well
// (mesher.hpp), so that neither the input nor the output has to fit in host memory.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <csignal>
#include <sys/resource.h>
#include <dirent.h>
#include <execinfo.h>
#include <fstream>
#include <future>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <vector>

#include "bucket.hpp"
#include "mesher.hpp"
#include "mlsb200.hpp"
#include "ply_io.hpp"
#include "statistics.hpp"
#include "stream.hpp"

namespace
{

namespace st = mlsb::statistics;

struct Options
{
    // src/mlsgpu_core.cpp:100-139 (defaults), :73-98 (names)
    double fitSmooth = 4.0;
    double maxRadius = std::numeric_limits<float>::infinity();
    double fitGrid = 0.01;
    double fitPrune = 0.02;
    double fitBoundaryLimit = 1.0;
    std::string fitShape = "sphere";
    int levels = 6;
    int subsampling = 3;
    int leafCells = 63;
    int deviceThreads = 4;           // the reference defaults to 1 and recommends more (src/marching.h:573-576)
    std::uint64_t memBucketSplats = 64ull << 20;   // bytes
    // Inputs larger than this are streamed from disk batch by batch (stream.hpp).  The reference's default is
    // 256M (2013 hosts); with it every input above 8 M splats takes the multi-pass path, so the default here
    // is sized for the hosts B200s sit in.  Pass the reference's value to get its behaviour.
    std::uint64_t memLoadSplats = 16ull << 30;
    std::uint64_t memMesh = 512ull << 20;
    // The mesher can work out of core like the reference's (src/mesher.cpp:343-445): finished blocks go to a
    // temporary file in tmpDir (the reference's default is the system's temporary directory too) once memHostMesh
    // bytes of mesh are held in memory.  Measured on C3 (100 M splats, 16 GB of mesh) with EVERYTHING going through
    // the file: 5.7 instead of 38.4 GB resident; 11.8 against 14.7 s on one box and 15.4 against 11.0 s on another —
    // the mesh is written twice, so it follows the temporary directory's disk.  Hence the budget: a mesh that fits
    // stays in memory, one that does not still gets done.
    std::string tmpDir;
    bool inCore = false;
    std::int64_t memHostMesh = -1;   // < 0: a quarter of the physical memory
    std::vector<int> devices;        // CUDA ordinals; empty = every device found (the reference's default, src/clh.cpp:180-240)
    bool quiet = false, statistics = false, statisticsCL = false, split = false;
    std::uint64_t splitSize = 100ull << 20;   // src/mlsgpu_core.cpp:162
    std::string checkpoint, resume;           // src/mlsgpu_core.cpp:121-122
    std::string statisticsFile, timeplot;
    std::string output;
    std::vector<std::string> inputs;
};

void usage()
{
    std::cout <<
        "Usage: mlsgpu [options] -o output.ply input.ply [input.ply...]\n"
        "  -h, --help                  show this help\n"
        "  -q, --quiet                 do not show informational messages\n"
        "  -o, --output-file FILE      output file (required)\n"
        "  --fit-smooth X              smoothing factor (4)\n"
        "  --max-radius X              limit influence radii\n"
        "  --fit-grid X                spacing of grid cells (0.01)\n"
        "  --fit-prune X               minimum fraction of vertices per component (0.02)\n"
        "  --fit-boundary-limit X      tuning factor for boundary detection (1)\n"
        "  --fit-shape sphere|plane    model shape (sphere)\n"
        "  --levels N                  levels in octree (6)\n"
        "  --subsampling N             subsampling of octree (3)\n"
        "  --leaf-cells N              leaf size for initial histogram (63)\n"
        "  --device-threads N          worker threads per device (4)\n"
        "  --mem-bucket-splats SIZE    device memory for splats of one bin, K/M/G suffix (64M)\n"
        "  --mem-load-splats SIZE      host memory for splats loaded at once; larger inputs are streamed from\n"
        "                              disk in several passes (16G; the reference's default is 256M)\n"
        "  --mem-host-mesh SIZE        host memory for the mesh (a quarter of the physical memory): finished blocks stay\n"
        "                              in memory up to that size, the ones after that go to a temporary file (40 bytes\n"
        "                              per vertex) and the output is streamed from both; 0 = everything goes to the file\n"
        "  --tmp-dir DIR               directory for that temporary file ($TMPDIR, else /tmp)\n"
        "  --in-core                   never use a temporary file (implied by --checkpoint)\n"
        "  --mem-mesh SIZE             device memory for raw mesh data (512M)\n"
        "  --device N[,N...]|all       CUDA devices to use, one worker group each (replaces --cl-device; default: all)\n"
        "  --cl-gpu                    use every GPU (the default)\n"
        "  --split                     split the output into a grid of files BASE_XXXX_YYYY_ZZZZ.ply (-o gives BASE)\n"
        "  --split-size SIZE           approximate size of the output chunks (100M)\n"
        "  --checkpoint FILE           save the mesher's state to FILE instead of writing the output\n"
        "  --resume FILE               write the output from a checkpoint (no input is read, no device is used)\n"
        "  --statistics                print statistics at the end\n"
        "  --statistics-file FILE      direct statistics to FILE instead of stdout (implies --statistics)\n"
        "  --statistics-cl             collect timings for the device kernels (kernel.*.time)\n"
        "  --timeplot FILE             write an EVENT log of what every thread did when (utils/timeplot.py's input)\n";
}

// Capacity, src/options.cpp:38-69
std::uint64_t parseCapacity(const std::string &s)
{
    char *end = nullptr;
    double v = std::strtod(s.c_str(), &end);
    std::uint64_t mul = 1;
    if (*end == 'K' || *end == 'k') mul = 1ull << 10;
    else if (*end == 'M' || *end == 'm') mul = 1ull << 20;
    else if (*end == 'G' || *end == 'g') mul = 1ull << 30;
    else if (*end != '\0' && *end != 'B') throw std::invalid_argument("bad capacity " + s);
    return (std::uint64_t) (v * mul);
}

Options parse(int argc, char **argv)
{
    Options o;
    // accepted for compatibility, value ignored
    static const std::map<std::string, bool> ignored = {
        {"--response-file", true},
        {"--max-split", true}, {"--reader", true}, {"--writer", true}, {"--omp-threads", true},
        {"--decache", false},
        {"--mem-host-splats", true}, {"--mem-reorder", true}, {"--mem-gather", true},
        {"--cl-device", true}, {"--cl-gpu", false}, {"--cl-cpu", false}, {"--debug", false}};
    for (int i = 1; i < argc; i++)
    {
        std::string a = argv[i], v;
        auto eq = a.find('=');
        bool hasValue = false;
        if (a.rfind("--", 0) == 0 && eq != std::string::npos)
        {
            v = a.substr(eq + 1);
            a = a.substr(0, eq);
            hasValue = true;
        }
        auto value = [&]() -> std::string {
            if (hasValue) return v;
            if (i + 1 >= argc) throw std::invalid_argument("option " + a + " needs a value");
            return argv[++i];
        };
        if (a == "-h" || a == "--help") { usage(); std::exit(0); }
        else if (a == "-q" || a == "--quiet") o.quiet = true;
        else if (a == "-o" || a == "--output-file") o.output = value();
        else if (a == "--fit-smooth") o.fitSmooth = std::stod(value());
        else if (a == "--max-radius") o.maxRadius = std::stod(value());
        else if (a == "--fit-grid") o.fitGrid = std::stod(value());
        else if (a == "--fit-prune") o.fitPrune = std::stod(value());
        else if (a == "--fit-boundary-limit") o.fitBoundaryLimit = std::stod(value());
        else if (a == "--fit-shape") o.fitShape = value();
        else if (a == "--levels") o.levels = std::stoi(value());
        else if (a == "--subsampling") o.subsampling = std::stoi(value());
        else if (a == "--leaf-cells") o.leafCells = std::stoi(value());
        else if (a == "--device-threads") o.deviceThreads = std::stoi(value());
        else if (a == "--mem-bucket-splats") o.memBucketSplats = parseCapacity(value());
        else if (a == "--mem-mesh") o.memMesh = parseCapacity(value());
        else if (a == "--mem-load-splats") o.memLoadSplats = parseCapacity(value());
        else if (a == "--device" || a == "--devices")
        {
            // --cl-device name[:n] selects among OpenCL devices (src/clh.cpp:170-178); CUDA has ordinals
            const std::string list = value();
            o.devices.clear();
            if (list != "all")
            {
                std::stringstream ss(list);
                std::string item;
                while (std::getline(ss, item, ','))
                    if (!item.empty()) o.devices.push_back(std::stoi(item));
                if (o.devices.empty()) throw std::invalid_argument("bad device list " + list);
            }
        }
        else if (a == "--statistics") o.statistics = true;
        else if (a == "--statistics-file") { o.statisticsFile = value(); o.statistics = true; }
        else if (a == "--statistics-cl") o.statisticsCL = true;
        else if (a == "--timeplot") o.timeplot = value();
        else if (a == "--tmp-dir") o.tmpDir = value();
        else if (a == "--in-core") o.inCore = true;
        else if (a == "--mem-host-mesh") o.memHostMesh = (std::int64_t) parseCapacity(value());
        else if (a == "--checkpoint") o.checkpoint = value();
        else if (a == "--resume") o.resume = value();
        else if (a == "--split") o.split = true;
        else if (a == "--split-size") o.splitSize = parseCapacity(value());
        else if (ignored.count(a)) { if (ignored.at(a)) (void) value(); }
        else if (a.rfind("-", 0) == 0) throw std::invalid_argument("unknown option " + a);
        else o.inputs.push_back(a);
    }
    // validateOptions, src/mlsgpu_core.cpp:398-457
    if (o.output.empty()) throw std::invalid_argument("the option '--output-file' is required but missing");
    if (o.inputs.empty() && o.resume.empty()) throw std::invalid_argument("At least one input file must be specified.");
    if (o.levels < 1 || o.levels > 10) throw std::invalid_argument("Value of --levels must be in the range 1 to 10.");
    if (o.subsampling < 3) throw std::invalid_argument("Value of --subsampling must be at least 3.");
    if (o.levels + o.subsampling > 14) throw std::invalid_argument("Sum of --subsampling and --levels is too large.");
    if (o.levels + o.subsampling - 1 > 9) throw std::invalid_argument("bins above 512 corners per side are not supported by this build.");
    if (o.leafCells < 1) throw std::invalid_argument("Value of --leaf-cells must be positive.");
    if (o.deviceThreads < 1) throw std::invalid_argument("Value of --device-threads must be at least 1.");
    if (!(o.fitPrune >= 0.0 && o.fitPrune <= 1.0)) throw std::invalid_argument("Value of --fit-prune must be in [0, 1].");
    if (o.fitShape != "sphere" && o.fitShape != "plane") throw std::invalid_argument("Value of --fit-shape must be sphere or plane.");
    if (!(o.fitGrid > 0.0)) throw std::invalid_argument("Value of --fit-grid must be positive.");
    if (o.memBucketSplats < sizeof(mlsb_splat)) throw std::invalid_argument("Value of --mem-bucket-splats is too small.");
    if (o.memLoadSplats < o.memBucketSplats)   // src/mlsgpu_core.cpp:436-440
        throw std::invalid_argument("Value of --mem-load-splats must be at least that of --mem-bucket-splats.");
    if (!o.checkpoint.empty()) o.inCore = true;   // a checkpoint IS the mesh in one file: it is saved from memory
    if (o.tmpDir.empty())
    {
        const char *env = std::getenv("TMPDIR");
        o.tmpDir = env && *env ? env : "/tmp";
    }
    if (o.memHostMesh < 0)
    {
        const long pages = sysconf(_SC_PHYS_PAGES), pageSize = sysconf(_SC_PAGE_SIZE);
        o.memHostMesh = pages > 0 && pageSize > 0 ? (std::int64_t) pages / 4 * pageSize : (std::int64_t) 8 << 30;
    }
    return o;
}

void collectInputs(const std::vector<std::string> &args, std::vector<std::string> &files)
{
    for (const std::string &a : args)
    {
        struct stat st;
        if (stat(a.c_str(), &st) == 0 && S_ISDIR(st.st_mode))
        {
            // directories are scanned for *.ply, src/mlsgpu_core.cpp:540-566
            std::vector<std::string> found;
            if (DIR *d = opendir(a.c_str()))
            {
                while (dirent *e = readdir(d))
                {
                    std::string n = e->d_name;
                    if (n.size() > 4 && n.substr(n.size() - 4) == ".ply") found.push_back(a + "/" + n);
                }
                closedir(d);
            }
            std::sort(found.begin(), found.end());
            files.insert(files.end(), found.begin(), found.end());
        }
        else
            files.push_back(a);
    }
}

double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Written
{
    mlsb::Mesher::Result totals;   // component counts; vertices and triangles summed over the files
    std::size_t files = 0;
};

// OOCMesher::write (src/mesher.cpp:763-852): prune, then one file — or, for a mesher in --split
// mode, one file per chunk that holds at least one triangle, named by ChunkNamer (:88-96).
Written writeMesh(mlsb::Mesher &mesher, bool split, const std::string &output, double fitPrune, const mlsb::PlyWriter &writer)
{
    Written w;
    if (mesher.spilling())
    {
        // out of core: pruned, compacted and written block by block from the temporary file
        const std::vector<mlsb::Mesher::Result> chunks = mesher.stream(fitPrune,
            [&](const mlsb::Mesher::chunk_type &chunk, std::size_t numVertices, std::size_t numTriangles) {
                if (split && numTriangles == 0) return std::unique_ptr<mlsb::PlyStreamWriter>();
                w.files++;
                return writer.openStream(split ? mlsb::chunkName(output, chunk) : output, numVertices, numTriangles);
            });
        for (const mlsb::Mesher::Result &c : chunks)
        {
            w.totals.totalComponents = c.totalComponents;
            w.totals.keptComponents = c.keptComponents;
            w.totals.thresholdVertices = c.thresholdVertices;
            w.totals.numVertices += c.numVertices;
            w.totals.numTriangles += c.numTriangles;
        }
        if (!split && w.files == 0)   // no block at all: still an (empty) output file
        {
            writer.write(output, std::vector<mlsb::PlyWriter::vertex_type>(), std::vector<mlsb::PlyWriter::triangle_type>());
            w.files = 1;
        }
        return w;
    }
    if (!split)
    {
        mlsb::Mesher::Result mesh = mesher.finish(fitPrune);
        writer.write(output, mesh.vertexChunks, mesh.triangleChunks);
        w.totals.totalComponents = mesh.totalComponents;
        w.totals.keptComponents = mesh.keptComponents;
        w.totals.thresholdVertices = mesh.thresholdVertices;
        w.totals.numVertices = mesh.numVertices;
        w.totals.numTriangles = mesh.numTriangles;
        w.files = 1;
        return w;
    }
    const std::vector<mlsb::Mesher::Result> chunks = mesher.finishChunks(fitPrune);
    for (const mlsb::Mesher::Result &c : chunks)
    {
        w.totals.totalComponents = c.totalComponents;
        w.totals.keptComponents = c.keptComponents;
        w.totals.thresholdVertices = c.thresholdVertices;
        w.totals.numVertices += c.numVertices;
        w.totals.numTriangles += c.numTriangles;
        if (c.numTriangles == 0) continue;
        writer.write(mlsb::chunkName(output, c.chunk), c.vertexChunks, c.triangleChunks);
        w.files++;
    }
    return w;
}

// The mesher's summary, under the reference's names and kinds (src/mesher.cpp:530-536, :850)
void recordMeshStatistics(const Written &w, std::size_t verticesBeforePruning)
{
    const mlsb::Mesher::Result &m = w.totals;
    st::getStatistic<st::Variable>("components.vertices.threshold").add((double) m.thresholdVertices);
    st::getStatistic<st::Variable>("components.vertices.total").add((double) verticesBeforePruning);
    st::getStatistic<st::Variable>("components.vertices.kept").add((double) m.numVertices);
    st::getStatistic<st::Variable>("components.triangles.kept").add((double) m.numTriangles);
    st::getStatistic<st::Variable>("components.total").add((double) m.totalComponents);
    st::getStatistic<st::Variable>("components.kept").add((double) m.keptComponents);
    st::getStatistic<st::Counter>("output.files").add(w.files);
    st::getStatistic<st::Counter>("output.vertices").add(m.numVertices);
    st::getStatistic<st::Counter>("output.triangles").add(m.numTriangles);
}

// writeStatistics (src/mlsgpu_core.cpp:313-352): version, options, then the registry in name order
// with 15 significant digits, to --statistics-file or stdout.
void writeStatistics(const std::string &file, int argc, char **argv)
{
    std::ostringstream out;
    out << "mlsgpu version: " << mlsb_version() << "\n";
    out << "mlsgpu variant: mlsb200 sm_100a\n";
    out << "mlsgpu options:";
    for (int i = 1; i < argc; i++) out << ' ' << argv[i];
    out << "\n";
    out.precision(15);
    out << st::Registry::getInstance();
    if (!file.empty())
    {
        std::ofstream f(file.c_str());
        f << out.str();
        f.close();
        if (!f) throw std::runtime_error("Could not write statistics to " + file);
    }
    else
        std::cout << out.str();
}

// --timeplot (src/timeplot.cpp:134-186): "EVENT worker action t0 t1" lines, seconds on one clock
struct TimeplotLog
{
    bool enabled = false;
    std::mutex mutex;
    std::vector<std::string> lines;
    void add(const std::string &worker, const char *action, double t0, double t1)
    {
        if (!enabled) return;
        char buf[256];
        std::snprintf(buf, sizeof(buf), "EVENT %s %s %.6f %.6f", worker.c_str(), action, t0, t1);
        std::lock_guard<std::mutex> lock(mutex);
        lines.push_back(buf);
    }
};

struct SinkState
{
    explicit SinkState(bool split) : mesher(split) {}
    TimeplotLog timeplot;
    mlsb::Mesher mesher;
    std::mutex mutex;
    double prepareSeconds = 0.0;   // per-block work, concurrently on the worker threads (summed)
    double commitSeconds = 0.0;    // seam bookkeeping in the global mesh, serialised
};

// Called concurrently by the device workers (mlsb_group_set_concurrent_sink): the per-block part of
// the mesher runs in parallel, only the merge into the global mesh takes the lock.
void meshSink(std::uint64_t tag, const mlsb_mesh_sizes *sizes, const void *blob, void *user)
{
    SinkState *st = static_cast<SinkState *>(user);
    const double tSink = st->timeplot.enabled ? mlsb_timeplot_now() : 0.0;
    // the tag carries the bin's output chunk (ChunkId::coords), 21 bits per axis
    const mlsb::Mesher::chunk_type chunk = {{(std::uint32_t) (tag & 0x1FFFFF), (std::uint32_t) ((tag >> 21) & 0x1FFFFF),
                                             (std::uint32_t) ((tag >> 42) & 0x1FFFFF)}};
    const double t0 = now();
    std::unique_ptr<mlsb::Mesher::Block> prepared = mlsb::Mesher::prepare(*sizes, blob);
    const double t1 = now();
    mlsb::Mesher::Block *block;
    double commit;
    {
        std::lock_guard<std::mutex> lock(st->mutex);
        const double t2 = now();
        block = &st->mesher.commit(std::move(prepared), chunk);
        commit = now() - t2;
        st->commitSeconds += commit;
    }
    const double t3 = now();
    mlsb::Mesher::complete(*block);
    st->mesher.spill(*block);   // out-of-core mode (--tmp-dir): the block's arrays leave host memory now
    const double local = (t1 - t0) + (now() - t3);
    {
        std::lock_guard<std::mutex> lock(st->mutex);
        st->prepareSeconds += local;
    }
    if (st->timeplot.enabled) st->timeplot.add("mesher", "compute", tSink, mlsb_timeplot_now());   // MesherGroup's worker
}



// A crash report that survives being run under a harness that swallows core dumps: the raw return
// addresses (resolve with addr2line -e mlsgpu) go to stderr before the default action runs.
void crashHandler(int sig)
{
    void *frames[64];
    const int n = backtrace(frames, 64);
    const char msg[] = "mlsgpu: fatal signal, backtrace follows\n";
    if (write(2, msg, sizeof(msg) - 1) < 0) {}
    backtrace_symbols_fd(frames, n, 2);
    signal(sig, SIG_DFL);
    raise(sig);
}

}  // namespace

int main(int argc, char **argv)
{
    signal(SIGSEGV, crashHandler);
    signal(SIGBUS, crashHandler);
    signal(SIGABRT, crashHandler);
    try
    {
        const Options o = parse(argc, argv);
        const double t0 = now();

        if (!o.resume.empty())
        {
            // --resume: the mesher's state comes from a checkpoint; only pruning and writing are left
            // (src/mlsgpu.cpp:105-109).  --fit-prune applies now; the chunking is the checkpoint's.
            std::ifstream file(o.resume.c_str(), std::ios::binary);
            if (!file) throw std::runtime_error("Could not open " + o.resume);
            mlsb::Mesher mesher;
            mesher.load(file);
            mlsb::PlyWriter writer;
            writer.addComment("mlsgpu (mlsb200) resumed from " + o.resume);
            const std::size_t before = mesher.numVertices();
            const Written written = writeMesh(mesher, mesher.isSplit(), o.output, o.fitPrune, writer);
            recordMeshStatistics(written, before);
            st::getStatistic<st::Variable>("run.time").add(now() - t0);
            if (!o.quiet || o.statistics) writeStatistics(o.statisticsFile, argc, argv);
            return 0;
        }

        // ---- the inputs are opened (headers parsed) before any device is touched: a bad file name or a
        // malformed PLY is reported at once, as the reference does when it builds its FileSet ----
        std::vector<std::string> files;
        collectInputs(o.inputs, files);
        std::vector<std::unique_ptr<mlsb::PlyReader>> readers;
        std::vector<std::size_t> firsts(1, 0);
        for (const std::string &f : files)
        {
            readers.emplace_back(new mlsb::PlyReader(f, (float) o.fitSmooth, (float) o.maxRadius));
            firsts.push_back(firsts.back() + readers.back()->size());
        }
        {
            // src/mlsgpu_core.cpp:568-570
            std::uint64_t totalBytes = 0;
            for (const std::unique_ptr<mlsb::PlyReader> &r : readers) totalBytes += (std::uint64_t) r->size() * r->vertexBytes();
            st::getStatistic<st::Counter>("files.scans").add(files.size());
            st::getStatistic<st::Counter>("files.splats").add(firsts.back());
            st::getStatistic<st::Counter>("files.bytes").add(totalBytes);
        }

        // ---- device workers: one group per device (the reference: one DeviceWorkerGroup per device,
        // src/mlsgpu_core.cpp:704-741), created in the background while the input is read and bucketed.
        // The buffers are sized from the options alone (--mem-bucket-splats, --levels,
        // --subsampling), as DeviceWorkerGroup's are (src/mlsgpu_core.cpp:722-735). ----
        const std::int64_t maxCells = (std::int64_t(1) << (o.levels + o.subsampling - 1)) - 1;   // src/mlsgpu_core.cpp:722
        const std::size_t maxBucketSplats = (std::size_t) (o.memBucketSplats / sizeof(mlsb_splat));
        std::vector<int> devices = o.devices;
        if (devices.empty())
        {
            int n = 0;
            mlsb::check(mlsb_device_count(&n));
            if (n == 0) throw std::runtime_error("No suitable CUDA device found");   // "No suitable OpenCL device found", mlsgpu.cpp:221
            for (int d = 0; d < n; d++) devices.push_back(d);
        }
        const std::size_t numGroups = devices.size();
        std::vector<mlsb_worker_config> cfgs(numGroups);
        for (std::size_t g = 0; g < numGroups; g++)
        {
            mlsb_worker_config &cfg = cfgs[g];
            cfg.device = devices[g];
            cfg.maxLevels = (std::uint32_t) o.levels;
            cfg.subsampling = (std::uint32_t) o.subsampling;
            cfg.maxSplats = (std::uint32_t) std::max<std::size_t>(maxBucketSplats, 1);
            cfg.maxCells = (std::uint32_t) maxCells;
            cfg.meshBytes = std::max<std::uint64_t>(o.memMesh / (o.deviceThreads + 1), (std::uint64_t) maxCells * maxCells * 872);
            cfg.shape = o.fitShape == "plane" ? MLSB_SHAPE_PLANE : MLSB_SHAPE_SPHERE;
            cfg.boundaryLimit = (float) o.fitBoundaryLimit;
        }
        // Everything the worker threads and the sink read is declared BEFORE the guard that joins them,
        // so that on any exit path (an exception included) it outlives the workers.
        for (std::size_t g = 0; g < numGroups; g++)
        {
            // CLH::ResourceUsage::addStatistics (src/clh.cpp:156-162) for what the groups will allocate
            mlsb_resource_usage usage;
            mlsb::check(mlsb_worker_resource_usage_ex(&cfgs[g], &usage));
            char name[64];
            std::snprintf(name, sizeof(name), "mem.device.%d", devices[g]);
            st::getStatistic<st::Peak>(name) += (std::int64_t) (usage.deviceBytes * (std::uint64_t) o.deviceThreads);
            st::getStatistic<st::Peak>("mem.device.all") += (std::int64_t) (usage.deviceBytes * (std::uint64_t) o.deviceThreads);
            st::getStatistic<st::Peak>("mem.pinned") += (std::int64_t) (usage.pinnedHostBytes * (std::uint64_t) o.deviceThreads);
        }
        SinkState state(o.split);
        if (!o.inCore) state.mesher.enableSpill(o.tmpDir, (std::uint64_t) o.memHostMesh);
        state.timeplot.enabled = !o.timeplot.empty();
        const double tp0 = mlsb_timeplot_now();
        std::vector<mlsb_splat> splats;
        std::vector<mlsb::Bin> bins;
        std::vector<std::unique_ptr<mlsb_splat[]>> ring;
        struct Created
        {
            mlsb_group *group = nullptr;
            int status = MLSB_OK;
            std::string message;
            double seconds = 0.0;
        };
        const bool streaming = (std::uint64_t) firsts.back() * sizeof(mlsb_splat) > o.memLoadSplats;
        if (!streaming)
        {
            // allocated and touched (in parallel) before the creation threads start, see the read below
            splats.resize(firsts.back());
        }
        std::vector<std::future<Created>> creating;
        for (std::size_t g = 0; g < numGroups; g++)
            creating.push_back(std::async(std::launch::async, [&cfgs, &o, &state, g] {
                Created c;
                const double t = now();
                c.status = mlsb_group_create(&cfgs[g], (std::uint32_t) o.deviceThreads, 0, 1, meshSink, &state, &c.group);
                if (c.status != MLSB_OK) c.message = mlsb_last_error_string();   // the error string is per thread
                c.seconds = now() - t;
                return c;
            }));
        // whatever happens below, the groups are joined and destroyed before anything above goes away
        struct GroupGuard
        {
            std::vector<std::future<Created>> &creating;
            std::vector<mlsb_group *> groups;
            explicit GroupGuard(std::vector<std::future<Created>> &c) : creating(c) {}
            ~GroupGuard()
            {
                for (std::future<Created> &f : creating)
                    if (f.valid()) groups.push_back(f.get().group);
                for (mlsb_group *g : groups) mlsb_group_destroy(g);
            }
        } guard(creating);

        // ---- grid and bins, in core or streamed (doComputeBlobs / doBucket) ----
        const std::int64_t bucketSize = std::min<std::int64_t>(o.leafCells, maxCells);             // :600-603
        // --split: chunks of about --split-size bytes, assuming a surface cuts a chunk of x cells per
        // side in ~20 x^2 vertices of 38 bytes each (src/mlsgpu_core.cpp:635-652)
        std::int64_t chunkCells = 0;
        if (o.split) chunkCells = std::max<std::int64_t>(1, (std::int64_t) std::ceil(std::sqrt(o.splitSize / 760.0)));
        mlsb::SplatStream stream(readers);
        mlsb::Grid grid;
        std::size_t numSplats = 0;
        if (!streaming)
        {
            // parallel decode, chunk by chunk with pread into the pre-faulted array (no new mappings are made
            // while the devices' worker groups are being created: on an 8-GPU box the creation threads pin
            // gigabytes of host memory, and a reader that page-faults its way through an mmap of the input
            // stalls behind them — 8 s instead of 1.7 s for 100 M splats); non-finite splats are dropped
            // afterwards (src/splat_set.cpp:221-240), keeping the order
            struct Piece { std::size_t file, first, count, dst; };
            std::vector<Piece> pieces;
            for (std::size_t f = 0; f < readers.size(); f++)
                for (std::size_t first = 0; first < readers[f]->size(); first += mlsb::SplatStream::kChunkVertices)
                    pieces.push_back(Piece{f, first, std::min(mlsb::SplatStream::kChunkVertices, readers[f]->size() - first), firsts[f] + first});
#pragma omp parallel
            {
                std::vector<char> scratch;
#pragma omp for schedule(dynamic, 4)
                for (std::int64_t i = 0; i < (std::int64_t) pieces.size(); i++)
                {
                    const Piece &p = pieces[(std::size_t) i];
                    readers[p.file]->readRange(p.first, p.count, splats.data() + p.dst, scratch);
                }
            }
            std::size_t kept = 0;
            for (std::size_t i = 0; i < splats.size(); i++)
                if (mlsb::PlyReader::isFinite(splats[i]))
                {
                    if (kept != i) splats[kept] = splats[i];
                    kept++;
                }
            splats.resize(kept);
            numSplats = kept;
        }
        else
            numSplats = stream.scan();   // pass 1: bounds and counts, one chunk per thread at a time
        if (numSplats == 0) throw std::runtime_error("At least one input point is required.");
        st::getStatistic<st::Counter>("splats.nonfinite").add(firsts.back() - numSplats);
        st::Peak &memSplats = st::getStatistic<st::Peak>("mem.splats");   // host bytes of the splats loaded at once
        memSplats = (std::int64_t) (splats.capacity() * sizeof(mlsb_splat));
        const double tRead = now();
        grid = streaming ? stream.grid((float) o.fitGrid, bucketSize) : mlsb::boundingGrid(splats, (float) o.fitGrid, bucketSize);
        for (int a = 0; a < 3; a++)
            if (grid.numCells(a) + 1 > (1 << 20) - 1)   // src/mlsgpu_core.cpp:623-629
                throw std::runtime_error("The bounding box is too big. Perhaps you have used the wrong units for --fit-grid.");
        for (int a = 0; a < 3; a++)   // postprocessGrid, src/mlsgpu_core.cpp:618-622
            st::getStatistic<st::Variable>(std::string("bbox") + "XYZ"[a]).add((double) grid.numCells(a) * grid.spacing);
        const mlsb::BlockLayout layout(grid, maxCells, chunkCells);
        std::vector<std::pair<std::size_t, std::size_t>> batches;   // streaming: runs of blocks loaded together
        if (!streaming)
        {
            mlsb::worldToGrid(splats, grid);
            bins = mlsb::makeBins(splats, grid, maxCells, maxBucketSplats, chunkCells);
        }
        else
        {
            const std::vector<std::uint64_t> counts = stream.countBlocks(grid, layout);   // pass 2
            batches = mlsb::SplatStream::batches(counts, o.memLoadSplats / sizeof(mlsb_splat));
        }
        const double tBucket = now();

        // ---- device pipeline ----
        const double tCreate0 = now();
        std::vector<mlsb_group *> groups;
        double createSeconds = 0.0;
        {
            std::string failure;
            for (std::future<Created> &f : creating)
            {
                const Created c = f.get();
                guard.groups.push_back(c.group);
                groups.push_back(c.group);
                createSeconds = std::max(createSeconds, c.seconds);
                if (c.status != MLSB_OK && failure.empty()) failure = c.message;
            }
            if (!failure.empty()) throw std::runtime_error(failure);
        }
        float origin[3];
        grid.getVertex0(origin);
        state.timeplot.add("main", "load", tp0, mlsb_timeplot_now());       // read + grid + bins: the loader's part
        for (mlsb_group *group : groups)
        {
            mlsb::check(mlsb_group_set_concurrent_sink(group, 1));
            if (state.timeplot.enabled) mlsb::check(mlsb_group_set_timeplot(group, 1));
            for (int i = 0; i < o.deviceThreads; i++)
            {
                mlsb_worker *w = nullptr;
                mlsb::check(mlsb_group_worker(group, (std::uint32_t) i, &w));
                mlsb::check(mlsb_set_scale_bias(w, grid.spacing, origin));   // ScaleBiasFilter::setScaleBias(grid)
                if (o.statisticsCL) mlsb::check(mlsb_worker_set_profiling(w, 2));   // src/mlsgpu_core.cpp:225-228
            }
        }
        const double tCreateWait = now() - tCreate0;   // the part of the creation that was not hidden
        double tGather = 0.0, tLoad = 0.0;
        // Gather buffers are recycled (BucketLoader's circular buffer): a slot is reused once the
        // group it went to has released the bin that used it, which keeps the gather out of the
        // page-fault path.
        const std::size_t queueDepth = 2 * (std::size_t) o.deviceThreads;   // mlsb_group_create's default
        ring.resize(numGroups * (queueDepth + (std::size_t) o.deviceThreads) + 1);
        struct SlotUse { std::size_t group = 0; std::uint64_t sequence = 0; std::size_t capacity = 0; bool used = false; };
        std::vector<SlotUse> slotUse(ring.size());
        std::vector<std::uint64_t> submittedBins(numGroups, 0), submittedSplats(numGroups, 0);
        std::size_t numBins = 0, chunksRead = 0;
        // BucketLoader::operator() for a list of bins over `source` (src/bucket_loader.cpp:104-126)
        st::Variable &loaderWrite = st::getStatistic<st::Variable>("bucket.loader.write");
        st::Variable &copySplats = st::getStatistic<st::Variable>("copy.splats"), &copySize = st::getStatistic<st::Variable>("copy.size");
        st::Peak &memRing = st::getStatistic<st::Peak>("mem.ring");   // host bytes of the gather buffers in flight
        auto dispatch = [&](const std::vector<mlsb_splat> &source, const std::vector<mlsb::Bin> &list, std::size_t blockBegin, std::size_t blockEnd) {
            std::size_t collectedBins = 0, collectedSplats = 0;
            for (const mlsb::Bin &bin : list)
            {
                if (streaming)
                {
                    const std::size_t block = mlsb::blockOfBin(layout, bin);
                    if (block < blockBegin || block >= blockEnd) continue;   // incomplete here: another batch owns it
                }
                const double tg = now();
                const std::vector<std::uint32_t> &ids = bin.ids;
                const std::size_t slotIndex = numBins % ring.size();
                SlotUse &use = slotUse[slotIndex];
                if (use.used)
                    for (std::uint64_t released = 0;;)
                    {
                        mlsb::check(mlsb_group_released(groups[use.group], &released));
                        if (released > use.sequence) break;
                        std::this_thread::yield();
                    }
                if (use.capacity < ids.size())
                {
                    memRing -= (std::int64_t) (use.capacity * sizeof(mlsb_splat));
                    use.capacity = std::max(ids.size(), use.capacity + use.capacity / 2);
                    ring[slotIndex].reset(new mlsb_splat[use.capacity]);
                    memRing += (std::int64_t) (use.capacity * sizeof(mlsb_splat));
                }
                mlsb_splat *dst = ring[slotIndex].get();
                // a few threads are enough and leave the cores to the device workers
#pragma omp parallel for schedule(static) num_threads(4)
                for (std::int64_t i = 0; i < (std::int64_t) ids.size(); i++) dst[i] = source[ids[(std::size_t) i]];
                tGather += now() - tg;
                loaderWrite.add(now() - tg);                 // BucketLoader's "write" action, src/bucket_loader.cpp:109
                copySplats.add((double) ids.size());         // src/workers.cpp:414-415
                copySize.add((double) (bin.size[0] - 1) * (double) (bin.size[1] - 1) * (double) (bin.size[2] - 1));
                collectedBins++;
                collectedSplats += ids.size();
                // CopyGroupBase::Worker::flush (src/workers.cpp:329-343) hands a bin to the device with the most
                // unallocated splat capacity: here, the group with the fewest splats submitted and not yet processed
                std::size_t target = 0;
                std::uint64_t least = ~0ull;
                for (std::size_t g = 0; g < numGroups; g++)
                {
                    std::uint64_t doneSplats = 0;
                    mlsb::check(mlsb_group_counters(groups[g], nullptr, &doneSplats, nullptr));
                    const std::uint64_t pending = submittedSplats[g] - doneSplats;
                    if (pending < least) { least = pending; target = g; }
                }
                use.group = target;
                use.sequence = submittedBins[target]++;
                use.used = true;
                submittedSplats[target] += ids.size();
                const std::uint64_t tag = (std::uint64_t) bin.chunk[0] | ((std::uint64_t) bin.chunk[1] << 21) | ((std::uint64_t) bin.chunk[2] << 42);
                mlsb::check(mlsb_group_submit(groups[target], dst, (std::uint32_t) ids.size(), bin.size, bin.offset, tag));
                numBins++;
            }
            // one flush of the collector (src/bucket_collector.cpp:71-80): the bins handed over together
            st::getStatistic<st::Variable>("bucket.collector.bins").add((double) collectedBins);
            st::getStatistic<st::Variable>("bucket.collector.splats").add((double) collectedSplats);
        };
        if (!streaming)
            dispatch(splats, bins, 0, layout.numBlocks());
        else
            for (const std::pair<std::size_t, std::size_t> &batch : batches)
            {
                const double tl = now();
                std::size_t read = 0;
                stream.load(grid, layout, batch.first, batch.second, splats, &read);   // BucketCollector + BucketLoader
                chunksRead += read;
                const double tc = now();
                st::getStatistic<st::Variable>("bucket.loader.load").add(tc - tl);
                memSplats = (std::int64_t) (splats.capacity() * sizeof(mlsb_splat));
                bins = mlsb::makeBins(splats, grid, maxCells, maxBucketSplats, chunkCells);
                st::getStatistic<st::Variable>("bucket.loader.compute").add(now() - tc);
                tLoad += now() - tl;
                dispatch(splats, bins, batch.first, batch.second);   // gathered into the ring: splats is free again
            }
        {
            std::string failure;
            for (mlsb_group *group : groups)
                if (mlsb_group_wait(group) != MLSB_OK && failure.empty()) failure = mlsb_last_error_string();
            if (!failure.empty()) throw std::runtime_error(failure);
        }
        const double tDevice = now();
        const double tpDevice = mlsb_timeplot_now();

        // ---- mesher: prune, write (OOCMesher::write) ----
        mlsb::PlyWriter writer;
        {
            // options are echoed as comments, src/mlsgpu_core.cpp:680-685
            char buf[256];
            std::snprintf(buf, sizeof(buf), "mlsgpu (mlsb200) fit-smooth=%g fit-grid=%g fit-prune=%g fit-boundary-limit=%g fit-shape=%s levels=%d subsampling=%d",
                          o.fitSmooth, o.fitGrid, o.fitPrune, o.fitBoundaryLimit, o.fitShape.c_str(), o.levels, o.subsampling);
            writer.addComment(buf);
        }
        Written written;
        if (!o.checkpoint.empty())
        {
            // --checkpoint: keep the mesher's state instead of writing (src/mlsgpu.cpp:171-175)
            std::ofstream file(o.checkpoint.c_str(), std::ios::binary);
            if (!file) throw std::runtime_error("Could not open " + o.checkpoint);
            state.mesher.save(file);
        }
        else
        {
            const std::size_t before = state.mesher.numVertices();
            written = writeMesh(state.mesher, o.split, o.output, o.fitPrune, writer);
            recordMeshStatistics(written, before);
        }
        state.timeplot.add("main", "write", tpDevice, mlsb_timeplot_now());   // OOCMesher::write
        if (state.timeplot.enabled)
        {
            for (mlsb_group *group : groups)
                mlsb::check(mlsb_group_timeplot(group, [](void *user, const char *worker, const char *action, double t0, double t1) {
                    static_cast<TimeplotLog *>(user)->add(worker, action, t0, t1);
                }, &state.timeplot));
            std::ofstream file(o.timeplot.c_str());
            for (const std::string &line : state.timeplot.lines) file << line << "\n";
            if (!file) throw std::runtime_error("Could not write " + o.timeplot);
        }
        const double tEnd = now();

        {
            // ---- everything measured goes through the registry (src/statistics.h) ----
            st::getStatistic<st::Counter>("bins").add(numBins);
            st::getStatistic<st::Counter>("devices").add(numGroups);
            st::getStatistic<st::Variable>("read.time").add(tRead - t0);
            st::getStatistic<st::Variable>("bucket.time").add(tBucket - tRead);
            st::getStatistic<st::Variable>("device.time").add(tDevice - tBucket);
            st::getStatistic<st::Variable>("mesher.time").add(tEnd - tDevice);
            st::getStatistic<st::Variable>("run.time").add(tEnd - t0);
            st::getStatistic<st::Variable>("device.create.time").add(createSeconds);
            // the CUDA driver's start-up (no persistence daemon: the GPU is initialised per process) that
            // reading and bucketing did not hide
            st::getStatistic<st::Variable>("device.create.wait").add(tCreateWait);
            st::getStatistic<st::Variable>("device.gather.time").add(tGather);
            st::getStatistic<st::Variable>("mesher.prepare.time").add(state.prepareSeconds);
            st::getStatistic<st::Variable>("mesher.commit.time").add(state.commitSeconds);
            if (state.mesher.spilling()) st::getStatistic<st::Counter>("mesher.tmp.bytes").add(state.mesher.spilledBytes());
            {
                // the process's high-water mark of resident memory (pinned buffers and page tables included)
                struct rusage usage;
                if (getrusage(RUSAGE_SELF, &usage) == 0) st::getStatistic<st::Peak>("mem.rss") = (std::int64_t) usage.ru_maxrss * 1024;
            }
            if (streaming)
            {
                st::getStatistic<st::Counter>("stream.batches").add(batches.size());
                st::getStatistic<st::Counter>("stream.chunks").add(stream.numChunks());
                st::getStatistic<st::Counter>("stream.chunks.read").add(chunksRead);
                st::getStatistic<st::Variable>("stream.load.time").add(tLoad);
            }
            char name[64];
            for (std::size_t g = 0; g < numGroups; g++)
            {
                std::snprintf(name, sizeof(name), "device.%d.bins", devices[g]);
                st::getStatistic<st::Counter>(name).add(submittedBins[g]);
                std::snprintf(name, sizeof(name), "device.%d.splats", devices[g]);
                st::getStatistic<st::Counter>(name).add(submittedSplats[g]);
            }
            // the workers' own statistics: kernel.*.time with --statistics-cl (CUDA events under the reference's names,
            // src/statistics_cl.cpp), marching.shipouts and marching.overflow always (src/marching.cpp:350-352)
            for (mlsb_group *group : groups)
                for (int i = 0; i < o.deviceThreads; i++)
                {
                    mlsb_worker *w = nullptr;
                    mlsb::check(mlsb_group_worker(group, (std::uint32_t) i, &w));
                    mlsb::check(mlsb_worker_visit_statistics(w, [](void *, const char *statName, double sum, double sum2, std::uint64_t count) {
                        if (std::string(statName) == "marching.overflow")
                            st::getStatistic<st::Counter>(statName).add((unsigned long long) sum);
                        else
                            st::getStatistic<st::Variable>(statName).addSamples(sum, sum2, count);
                    }, nullptr));
                }
            st::getStatistic<st::Variable>("splats/s end-to-end").add(numSplats / (tEnd - t0));
            st::getStatistic<st::Variable>("splats/s excluding device start-up").add(numSplats / (tEnd - t0 - tCreateWait));
        }
        if (!o.quiet || o.statistics) writeStatistics(o.statisticsFile, argc, argv);
        return 0;
    }
    catch (const std::exception &e)
    {
        std::cerr << "mlsgpu: " << e.what() << "\n";
        return 1;
    }
}
