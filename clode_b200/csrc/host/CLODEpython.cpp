// pybind11 module `clode_cpp_wrapper` — the Python-visible surface of clode/cpp/CLODEpython.cpp:35-381
// (same class, method, enum and argument names; see also clode/cpp/clode_cpp_wrapper.pyi) over the
// B200 host classes.  Additions are marked: ndarray fast paths and measurement hooks.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "CLODE.hpp"
#include "CLODEfeatures.hpp"
#include "CLODEtrajectory.hpp"
#include "clode_log.hpp"

#include <cstdlib>
#include <map>
#include <mutex>

namespace py = pybind11;
using dvec = std::vector<double>;
using darray = py::array_t<double, py::array::c_style | py::array::forcecast>;

static std::string vector_to_string(const std::vector<std::string> &vec)
{
    std::string out = "[";
    for (const std::string &s : vec) out += s + ", ";
    return out + "]";
}

static py::array_t<double> to_array(const dvec &v)
{
    py::array_t<double> a((py::ssize_t)v.size());
    std::copy(v.begin(), v.end(), a.mutable_data());
    return a;
}

// Result arrays handed to numpy are backed by PAGE-LOCKED memory from the runtime (clode_host_alloc), so the
// device-to-host copy that fills them runs at the full PCIe rate and lands in the array the user gets — no std::vector
// in between (the reference: device -> std::vector -> Python list -> np.array, clode/features.py:516-524).  Pinning
// pages is slow (~0.3 ms/MB), so blocks go back to a small pool when numpy drops the array and are reused by size.
namespace {
struct PinnedPool {
    std::mutex m;
    std::multimap<size_t, void *> idle;
    size_t idle_bytes = 0;
    static constexpr size_t kMaxIdle = size_t(4) << 30;
    struct Block { void *p; size_t bytes; bool pinned; };

    Block take(size_t bytes)
    {
        bytes = std::max<size_t>((bytes + 4095) & ~size_t(4095), 4096);
        {
            std::lock_guard<std::mutex> lock(m);
            auto it = idle.find(bytes);
            if (it != idle.end()) {
                void *p = it->second;
                idle.erase(it);
                idle_bytes -= bytes;
                return {p, bytes, true};
            }
        }
        if (void *p = clode_host_alloc(0, bytes)) return {p, bytes, true};
        return {std::malloc(bytes), bytes, false}; // no driver (CPU-only import) or out of lockable memory: pageable
    }
    void give(Block b)
    {
        if (!b.pinned) { std::free(b.p); return; }
        std::lock_guard<std::mutex> lock(m);
        if (idle_bytes + b.bytes > kMaxIdle) { clode_host_free(b.p); return; }
        idle.emplace(b.bytes, b.p);
        idle_bytes += b.bytes;
    }
    static PinnedPool &get() { static PinnedPool *pool = new PinnedPool(); return *pool; } // leaked on purpose: outlives numpy arrays at exit
};

// a flat float64 array of `count` elements on a pooled page-locked block, filled by `fill(double *)`
template <class Fill> py::array_t<double> pinned_array(size_t count, Fill &&fill)
{
    PinnedPool::Block b = PinnedPool::get().take(std::max<size_t>(count, 1) * sizeof(double));
    if (!b.p) throw std::bad_alloc();
    try {
        py::gil_scoped_release release;
        fill(static_cast<double *>(b.p));
    } catch (...) {
        PinnedPool::get().give(b);
        throw;
    }
    auto *owner = new PinnedPool::Block(b);
    py::capsule keep(owner, [](void *o) {
        auto *blk = static_cast<PinnedPool::Block *>(o);
        PinnedPool::get().give(*blk);
        delete blk;
    });
    return py::array_t<double>({(py::ssize_t)count}, {(py::ssize_t)sizeof(double)}, static_cast<double *>(b.p), keep);
}

// (ensemble, cols) C-contiguous matrix, transposed on the GPU
template <class Fill> py::array_t<double> pinned_matrix(size_t n, size_t cols, Fill &&fill)
{
    py::array_t<double> flat = pinned_array(n * cols, fill);
    return flat.reshape({(py::ssize_t)n, (py::ssize_t)cols});
}
} // namespace

struct LoggerSingleton {
    LoggerSingleton()
    {
        auto &lg = clode_log::Logger::get();
        lg.level = clode_log::info;
        lg.sink = [](const std::string &line) {
            py::gil_scoped_acquire gil;
            py::print(line);
        };
    }
    static LoggerSingleton &instance()
    {
        static LoggerSingleton just_one;
        return just_one;
    }
    void set_log_level(clode_log::level_enum level) { clode_log::Logger::get().level = level; }
    void set_log_pattern(std::string &pattern) { clode_log::Logger::get().pattern = pattern; }
    clode_log::level_enum get_log_level() { return clode_log::Logger::get().level; }
};

PYBIND11_MODULE(clode_cpp_wrapper, m)
{
    m.doc() = "CLODE C++/Python interface (B200 runtime)";

    py::enum_<clode_log::level_enum>(m, "LogLevel")
        .value("trace", clode_log::trace)
        .value("debug", clode_log::debug)
        .value("info", clode_log::info)
        .value("warn", clode_log::warn)
        .value("err", clode_log::err)
        .value("critical", clode_log::critical)
        .value("off", clode_log::off)
        .export_values();

    py::class_<LoggerSingleton>(m, "LoggerSingleton")
        .def("set_log_level", &LoggerSingleton::set_log_level)
        .def("set_log_pattern", &LoggerSingleton::set_log_pattern)
        .def("get_log_level", &LoggerSingleton::get_log_level);
    m.def("get_logger", &LoggerSingleton::instance, py::return_value_policy::reference, "Get logger singleton instance");

    py::enum_<cl_vendor>(m, "CLVendor")
        .value("VENDOR_ANY", VENDOR_ANY)
        .value("VENDOR_NVIDIA", VENDOR_NVIDIA)
        .value("VENDOR_AMD", VENDOR_AMD)
        .value("VENDOR_INTEL", VENDOR_INTEL)
        .export_values();

    py::enum_<e_cl_device_type>(m, "CLDeviceType")
        .value("DEVICE_TYPE_ALL", DEVICE_TYPE_ALL)
        .value("DEVICE_TYPE_CPU", DEVICE_TYPE_CPU)
        .value("DEVICE_TYPE_GPU", DEVICE_TYPE_GPU)
        .value("DEVICE_TYPE_ACCELERATOR", DEVICE_TYPE_ACCELERATOR)
        .value("DEVICE_TYPE_DEFAULT", DEVICE_TYPE_DEFAULT)
        .value("DEVICE_TYPE_CUSTOM", DEVICE_TYPE_CUSTOM)
        .export_values();

    py::class_<OpenCLResource>(m, "OpenCLResource")
        .def(py::init<>())
        .def(py::init<cl_vendor>())
        .def(py::init<e_cl_device_type, cl_vendor>())
        .def(py::init<unsigned int, unsigned int>())
        .def(py::init<unsigned int, std::vector<unsigned int>>())
        .def("get_double_support", &OpenCLResource::getDoubleSupport, "Get double support", py::arg("device_id") = 0)
        .def("get_max_memory_alloc_size", &OpenCLResource::getMaxMemAllocSize, "Get max memory alloc size", py::arg("device_id") = 0)
        .def("get_device_cl_version", &OpenCLResource::getDeviceCLVersion, "Get device CL version", py::arg("device_id") = 0)
        .def("get_device_ordinals", &OpenCLResource::getDeviceOrdinals, "CUDA device indices of this resource (addition)")
        .def("print_devices", &OpenCLResource::print, "Print device info to log");

    py::class_<deviceInfo>(m, "DeviceInfo")
        .def_readwrite("name", &deviceInfo::name)
        .def_readwrite("vendor", &deviceInfo::vendor)
        .def_readwrite("version", &deviceInfo::version)
        .def_readwrite("device_type", &deviceInfo::devType)
        .def_readwrite("device_type_str", &deviceInfo::devTypeStr)
        .def_readwrite("compute_units", &deviceInfo::computeUnits)
        .def_readwrite("max_clock", &deviceInfo::maxClock)
        .def_readwrite("max_work_group_size", &deviceInfo::maxWorkGroupSize)
        .def_readwrite("device_memory_size", &deviceInfo::deviceMemSize)
        .def_readwrite("max_memory_alloc_size", &deviceInfo::maxMemAllocSize)
        .def_readwrite("extensions", &deviceInfo::extensions)
        .def_readwrite("double_support", &deviceInfo::doubleSupport)
        .def_readwrite("device_available", &deviceInfo::deviceAvailable)
        .def("__repr__", [](const deviceInfo &d) {
            return "<device_info(name=" + d.name + ", vendor=" + d.vendor + ", version=" + d.version +
                   ", device_type=" + d.devTypeStr + ", compute_units=" + std::to_string(d.computeUnits) +
                   ", max_clock=" + std::to_string(d.maxClock) + ", max_work_group_size=" + std::to_string(d.maxWorkGroupSize) +
                   ", device_memory_size=" + std::to_string(d.deviceMemSize) +
                   ", max_memory_alloc_size=" + std::to_string(d.maxMemAllocSize) + ", extensions=" + d.extensions +
                   ", double_support=" + std::to_string(d.doubleSupport) +
                   ", device_available=" + std::to_string(d.deviceAvailable) + ")>";
        });

    py::class_<platformInfo>(m, "PlatformInfo")
        .def_readwrite("name", &platformInfo::name)
        .def_readwrite("vendor", &platformInfo::vendor)
        .def_readwrite("version", &platformInfo::version)
        .def_readwrite("device_info", &platformInfo::device_info)
        .def_readwrite("device_count", &platformInfo::nDevices)
        .def("__repr__", [](const platformInfo &p) {
            return "<platform_info(name=" + p.name + ", vendor=" + p.vendor + ", version=" + p.version +
                   ", device_count=" + std::to_string(p.nDevices) + ")>";
        });

    m.def("query_opencl", &queryOpenCL, "Query OpenCL devices");
    m.def("_print_opencl", py::overload_cast<>(&printOpenCL), "Print OpenCL devices");

    py::class_<ProblemInfo>(m, "ProblemInfo")
        .def(py::init<const std::string &, const std::vector<std::string> &, const std::vector<std::string> &,
                      const std::vector<std::string> &, int>(),
             py::arg("src_file"), py::arg("vars"), py::arg("pars"), py::arg("aux") = std::vector<std::string>(),
             py::arg("num_noise") = 1)
        .def(py::init<>())
        .def_readwrite("src_file", &ProblemInfo::clRHSfilename)
        .def_readwrite("num_var", &ProblemInfo::nVar)
        .def_readwrite("num_par", &ProblemInfo::nPar)
        .def_readwrite("num_aux", &ProblemInfo::nAux)
        .def_readwrite("num_noise", &ProblemInfo::nWiener)
        .def_property("vars", &ProblemInfo::getVarNames, &ProblemInfo::setVarNames)
        .def_property("pars", &ProblemInfo::getParNames, &ProblemInfo::setParNames)
        .def_property("aux", &ProblemInfo::getAuxNames, &ProblemInfo::setAuxNames)
        .def("__repr__", [](const ProblemInfo &p) {
            return "<problem_info(src_file=" + p.clRHSfilename + ", vars=" + vector_to_string(p.varNames) +
                   ", pars=" + vector_to_string(p.parNames) + ", aux=" + vector_to_string(p.auxNames) +
                   ", num_noise=" + std::to_string(p.nWiener) + ")>";
        });

    py::class_<SolverParams<double>>(m, "SolverParams")
        .def(py::init([](double dt, double dtmax, double abstol, double reltol, unsigned int max_steps,
                         unsigned int max_store, unsigned int nout) {
                 return SolverParams<double>{dt, dtmax, abstol, reltol, max_steps, max_store, nout};
             }),
             py::arg("dt") = 0.1, py::arg("dtmax") = 0.5, py::arg("abstol") = 1e-6, py::arg("reltol") = 1e-3,
             py::arg("max_steps") = 1000000, py::arg("max_store") = 1000000, py::arg("nout") = 1)
        .def_readwrite("dt", &SolverParams<double>::dt)
        .def_readwrite("dtmax", &SolverParams<double>::dtmax)
        .def_readwrite("abstol", &SolverParams<double>::abstol)
        .def_readwrite("reltol", &SolverParams<double>::reltol)
        .def_readwrite("max_steps", &SolverParams<double>::max_steps)
        .def_readwrite("max_store", &SolverParams<double>::max_store)
        .def_readwrite("nout", &SolverParams<double>::nout)
        .def("__repr__", [](const SolverParams<double> &s) {
            return "<solver_params(dt=" + std::to_string(s.dt) + ", dtmax=" + std::to_string(s.dtmax) +
                   ", abstol=" + std::to_string(s.abstol) + ", reltol=" + std::to_string(s.reltol) +
                   ", max_steps=" + std::to_string(s.max_steps) + ", max_store=" + std::to_string(s.max_store) +
                   ", nout=" + std::to_string(s.nout) + ")>";
        });

    py::class_<CLODE>(m, "SimulatorBase")
        .def(py::init<ProblemInfo &, std::string &, bool, OpenCLResource &, std::string &>(), py::arg("problem_info"),
             py::arg("stepper"), py::arg("cl_single_precision"), py::arg("opencl_resource"), py::arg("clode_root"))
        .def("set_problem_info", &CLODE::setProblemInfo)
        .def("set_stepper", &CLODE::setStepper)
        .def("set_precision", &CLODE::setPrecision)
        .def("set_opencl", static_cast<void (CLODE::*)(OpenCLResource)>(&CLODE::setOpenCL))
        .def("set_opencl", static_cast<void (CLODE::*)(unsigned int, unsigned int)>(&CLODE::setOpenCL))
        .def("build_cl", &CLODE::buildCL)
        // ndarray fast paths first (one memcpy instead of a per-element list conversion), then the reference's list forms
        .def("set_problem_data", [](CLODE &c, const darray &x0, const darray &p) {
            const double *px = x0.data(), *pp = p.data();
            const size_t nx = (size_t)x0.size(), np_ = (size_t)p.size();
            py::gil_scoped_release release;
            c.setProblemData(px, nx, pp, np_);
        })
        .def("set_problem_data", static_cast<void (CLODE::*)(dvec, dvec)>(&CLODE::setProblemData))
        // additions: (ensemble, nVar) / (ensemble, nPar) float64 matrices with any positive strides, uploaded as they are
        .def("set_problem_data_matrix", [](CLODE &c, const py::array_t<double> &x0, const py::array_t<double> &p) {
            if (x0.ndim() != 2 || p.ndim() != 2) throw std::invalid_argument("set_problem_data_matrix: 2-D arrays expected");
            if (x0.shape(1) != c.getNvar() || p.shape(1) != c.getNpar()) throw std::invalid_argument("set_problem_data_matrix: wrong number of columns");
            const double *px = x0.data(), *pp = p.data();
            const size_t nx = (size_t)x0.shape(0), np_ = (size_t)p.shape(0);
            const ptrdiff_t xi = x0.strides(0) / 8, xv = x0.shape(1) > 1 ? x0.strides(1) / 8 : 1;
            const ptrdiff_t pi = p.strides(0) / 8, pv = p.shape(1) > 1 ? p.strides(1) / 8 : 1;
            py::gil_scoped_release release;
            c.setProblemDataMatrix(px, nx, nx > 1 ? xi : 1, xv, pp, np_, np_ > 1 ? pi : 1, pv);
        })
        .def("set_x0_matrix", [](CLODE &c, const py::array_t<double> &x0) {
            if (x0.ndim() != 2 || x0.shape(1) != c.getNvar()) throw std::invalid_argument("set_x0_matrix: (ensemble, nVar) array expected");
            c.setX0Matrix(x0.data(), (size_t)x0.shape(0), x0.shape(0) > 1 ? x0.strides(0) / 8 : 1, x0.shape(1) > 1 ? x0.strides(1) / 8 : 1);
        })
        .def("set_pars_matrix", [](CLODE &c, const py::array_t<double> &p) {
            if (p.ndim() != 2 || p.shape(1) != c.getNpar()) throw std::invalid_argument("set_pars_matrix: (ensemble, nPar) array expected");
            c.setParsMatrix(p.data(), (size_t)p.shape(0), p.shape(0) > 1 ? p.strides(0) / 8 : 1, p.shape(1) > 1 ? p.strides(1) / 8 : 1);
        })
        .def("set_tspan", &CLODE::setTspan)
        .def("set_x0", [](CLODE &c, const darray &x0) { c.setX0(x0.data(), (size_t)x0.size()); })
        .def("set_x0", static_cast<void (CLODE::*)(dvec)>(&CLODE::setX0))
        .def("set_pars", [](CLODE &c, const darray &p) { c.setPars(p.data(), (size_t)p.size()); })
        .def("set_pars", static_cast<void (CLODE::*)(dvec)>(&CLODE::setPars))
        .def("set_solver_params", &CLODE::setSolverParams)
        .def("seed_rng", static_cast<void (CLODE::*)()>(&CLODE::seedRNG), "Seed RNG")
        .def("seed_rng", static_cast<void (CLODE::*)(int)>(&CLODE::seedRNG), "Seed RNG", py::arg("seed"))
        .def("transient", &CLODE::transient, py::call_guard<py::gil_scoped_release>())
        .def("shift_tspan", &CLODE::shiftTspan)
        .def("shift_x0", &CLODE::shiftX0)
        .def("get_problem_info", &CLODE::getProblemInfo)
        .def("get_tspan", &CLODE::getTspan)
        .def("get_solver_params", &CLODE::getSolverParams)
        .def("get_pars", &CLODE::getPars)
        .def("get_x0", &CLODE::getX0)
        .def("get_xf", &CLODE::getXf)
        .def("get_dt", &CLODE::getDt)
        .def("get_tf", &CLODE::getTf)
        .def("get_available_steppers", &CLODE::getAvailableSteppers)
        .def("get_program_string", &CLODE::getProgramString)
        .def("print_status", &CLODE::printStatus)
        // additions
        .def("get_x0_array", [](CLODE &c) { return pinned_array((size_t)c.getNvar() * c.getNpts(), [&](double *o) { c.fetch(CLODE_BUF_X0, c.getNvar(), o, "CLODE::getX0"); }); })
        .def("get_xf_array", [](CLODE &c) { return pinned_array((size_t)c.getNvar() * c.getNpts(), [&](double *o) { c.fetch(CLODE_BUF_XF, c.getNvar(), o, "CLODE::getXf"); }); })
        .def("get_dt_array", [](CLODE &c) { return pinned_array((size_t)c.getNpts(), [&](double *o) { c.fetch(CLODE_BUF_DT, 1, o, "CLODE::getDt"); }); })
        .def("get_tf_array", [](CLODE &c) { return pinned_array((size_t)c.getNpts(), [&](double *o) { c.fetch(CLODE_BUF_TF, 1, o, "CLODE::getTf"); }); })
        .def("get_xf_matrix", [](CLODE &c) { return pinned_matrix((size_t)c.getNpts(), (size_t)c.getNvar(), [&](double *o) { c.fetchInstanceMajor(CLODE_BUF_XF, c.getNvar(), o, "CLODE::getXf"); }); })
        .def("get_x0_matrix", [](CLODE &c) { return pinned_matrix((size_t)c.getNpts(), (size_t)c.getNvar(), [&](double *o) { c.fetchInstanceMajor(CLODE_BUF_X0, c.getNvar(), o, "CLODE::getX0"); }); })
        .def("get_pars_array", [](CLODE &c) { return pinned_array((size_t)c.getNpar() * c.getNpts(), [&](double *o) { c.fetch(CLODE_BUF_PARS, c.getNpar(), o, "CLODE::getPars"); }); })
        .def("get_step_counts", [](CLODE &c) {
            auto v = c.getStepCounts();
            py::array_t<unsigned int> a((py::ssize_t)v.size());
            std::copy(v.begin(), v.end(), a.mutable_data());
            return a;
        })
        .def("get_rng_state", [](CLODE &c) {
            auto v = c.getRNGstate();
            py::array_t<std::uint64_t> a((py::ssize_t)v.size());
            std::copy(v.begin(), v.end(), a.mutable_data());
            return a;
        })
        .def("get_last_kernel_ms", &CLODE::getLastKernelMilliseconds);

    py::class_<ObserverParams<double>>(m, "ObserverParams")
        .def(py::init([](unsigned int e, unsigned int f, unsigned int mec, unsigned int met, double ma, double mi, double nr,
                         double xu, double xd, double dxu, double dxd, double eps) {
                 return ObserverParams<double>{e, f, mec, met, ma, mi, nr, xu, xd, dxu, dxd, eps};
             }),
             py::arg("e_var_ix") = 0, py::arg("f_var_ix") = 0, py::arg("max_event_count") = 100,
             py::arg("max_event_timestamps") = 0, py::arg("min_amp") = 0., py::arg("min_imi") = 0.,
             py::arg("nhood_radius") = 0.05, py::arg("x_up_threshold") = 0.2, py::arg("x_down_threshold") = 0.2,
             py::arg("dx_up_threshold") = 0., py::arg("dx_down_threshold") = 0., py::arg("eps_dx") = 0.)
        .def_readwrite("e_var_ix", &ObserverParams<double>::eVarIx)
        .def_readwrite("f_var_ix", &ObserverParams<double>::fVarIx)
        .def_readwrite("max_event_count", &ObserverParams<double>::maxEventCount)
        .def_readwrite("max_event_timestamps", &ObserverParams<double>::maxEventTimestamps)
        .def_readwrite("min_amp", &ObserverParams<double>::minXamp)
        .def_readwrite("min_imi", &ObserverParams<double>::minIMI)
        .def_readwrite("nhood_radius", &ObserverParams<double>::nHoodRadius)
        .def_readwrite("x_up_threshold", &ObserverParams<double>::xUpThresh)
        .def_readwrite("x_down_threshold", &ObserverParams<double>::xDownThresh)
        .def_readwrite("dx_up_threshold", &ObserverParams<double>::dxUpThresh)
        .def_readwrite("dx_down_threshold", &ObserverParams<double>::dxDownThresh)
        .def_readwrite("eps_dx", &ObserverParams<double>::eps_dx)
        .def("__repr__", [](const ObserverParams<double> &p) {
            return "<observer_params(e_var_ix=" + std::to_string(p.eVarIx) + ", f_var_ix=" + std::to_string(p.fVarIx) +
                   ", max_event_count=" + std::to_string(p.maxEventCount) +
                   ", max_event_timestamps=" + std::to_string(p.maxEventTimestamps) +
                   ", min_amp=" + std::to_string(p.minXamp) + ", min_imi=" + std::to_string(p.minIMI) +
                   ", nhood_radius=" + std::to_string(p.nHoodRadius) + ", x_up_threshold=" + std::to_string(p.xUpThresh) +
                   ", x_down_threshold=" + std::to_string(p.xDownThresh) + ", dx_up_threshold=" + std::to_string(p.dxUpThresh) +
                   ", dx_down_threshold=" + std::to_string(p.dxDownThresh) + ", eps_dx=" + std::to_string(p.eps_dx) + ")>";
        });

    py::class_<CLODEfeatures, CLODE>(m, "FeatureSimulatorBase")
        .def(py::init<ProblemInfo &, std::string &, std::string &, ObserverParams<double>, bool, OpenCLResource &, std::string &>())
        .def("build_cl", &CLODEfeatures::buildCL)
        .def("set_observer_params", &CLODEfeatures::setObserverParams)
        .def("set_observer", &CLODEfeatures::setObserver)
        .def("initialize_observer", &CLODEfeatures::initializeObserver, py::call_guard<py::gil_scoped_release>())
        .def("is_observer_initialized", &CLODEfeatures::isObserverInitialized)
        .def("features", static_cast<void (CLODEfeatures::*)(bool)>(&CLODEfeatures::features), py::call_guard<py::gil_scoped_release>())
        .def("features", static_cast<void (CLODEfeatures::*)()>(&CLODEfeatures::features), py::call_guard<py::gil_scoped_release>())
        .def("get_observer_params", &CLODEfeatures::getObserverParams)
        .def("get_observer_name", &CLODEfeatures::getObserverName)
        .def("get_f", &CLODEfeatures::getF)
        .def("get_f_array", [](CLODEfeatures &c) { return pinned_array((size_t)c.getNFeatures() * c.getNpts(), [&](double *o) { c.fetch(CLODE_BUF_F, c.getNFeatures(), o, "CLODEfeatures::getF"); }); })
        .def("get_f_matrix", [](CLODEfeatures &c) { return pinned_matrix((size_t)c.getNpts(), (size_t)c.getNFeatures(), [&](double *o) { c.fetchInstanceMajor(CLODE_BUF_F, c.getNFeatures(), o, "CLODEfeatures::getF"); }); })
        .def("get_n_features", &CLODEfeatures::getNFeatures)
        .def("get_feature_names", &CLODEfeatures::getFeatureNames)
        .def("get_available_observers", &CLODEfeatures::getAvailableObservers)
        .def("__repr__", [](const CLODEfeatures &c) {
            return "<CLODEfeatures (observer=" + c.getObserverName() + ", n_features=" + std::to_string(c.getNFeatures()) + ")>";
        });

    py::class_<CLODEtrajectory, CLODE>(m, "TrajectorySimulatorBase")
        .def(py::init<ProblemInfo &, std::string &, bool, OpenCLResource &, std::string &>())
        .def("build_cl", &CLODEtrajectory::buildCL)
        .def("trajectory", &CLODEtrajectory::trajectory, py::call_guard<py::gil_scoped_release>())
        .def("get_t", &CLODEtrajectory::getT)
        .def("get_x", &CLODEtrajectory::getX)
        .def("get_dx", &CLODEtrajectory::getDx)
        .def("get_aux", &CLODEtrajectory::getAux)
        .def("get_n_stored", &CLODEtrajectory::getNstored)
        .def("get_t_array", [](CLODEtrajectory &c) { return to_array(c.getT()); })
        .def("get_x_array", [](CLODEtrajectory &c) { return to_array(c.getX()); })
        .def("get_dx_array", [](CLODEtrajectory &c) { return to_array(c.getDx()); })
        .def("get_aux_array", [](CLODEtrajectory &c) { return to_array(c.getAux()); })
        .def("set_stream_chunk", &CLODEtrajectory::setStreamChunk, py::arg("rows"))
        .def("get_stream_chunk", &CLODEtrajectory::getStreamChunk);
}
