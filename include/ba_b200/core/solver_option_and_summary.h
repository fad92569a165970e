// Drop-in for core/solver_option_and_summary.h (reference :1-97) and .cpp (:1-87): Options, IterationStatus,
// OptimizationInfo and Summary with the same fields, defaults and BriefReport layout.  The reference's
// header includes "ceres/ceres.h" (:9) but uses nothing from it; it is not needed here.
#ifndef _SOLVER_OPTION_AND_SUMMARY_H_
#define _SOLVER_OPTION_AND_SUMMARY_H_
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace visual_navigation {
namespace analytic_solver {
#define TEXT_RED(str) (std::string("\033[0;31m") + str + std::string("\033[0m"))
#define TEXT_GREEN(str) (std::string("\033[0;32m") + str + std::string("\033[0m"))
#define TEXT_YELLOW(str) (std::string("\033[0;33m") + str + std::string("\033[0m"))
#define TEXT_BLUE(str) (std::string("\033[0;34m") + str + std::string("\033[0m"))
#define TEXT_MAGENTA(str) (std::string("\033[0;35m") + str + std::string("\033[0m"))
#define TEXT_CYAN(str) (std::string("\033[0;36m") + str + std::string("\033[0m"))

enum class SolverType { UNDEFINED = -1, GRADIENT_DESCENT = 0, GAUSS_NEWTON = 1, LEVENBERG_MARQUARDT = 2 };
enum class IterationStatus { UNDEFINED = -1, UPDATE = 0, UPDATE_TRUST_MORE = 1, SKIPPED = 2 };

struct OptimizationInfo {   // reference :37-46: one row per LM / GN iteration, every field starts at -1
  double cost{-1.0}, cost_change{-1.0}, average_reprojection_error{-1.0};
  double abs_gradient{-1.0}, abs_step{-1.0}, damping_term{-1.0}, iter_time{-1.0};
  IterationStatus iteration_status{IterationStatus::UNDEFINED};
};

class Options {   // reference :47-72: float thresholds / ratios on purpose (they are promoted in comparisons)
  friend class PoseOnlyBundleAdjustmentSolver;
  friend class FullBundleAdjustmentSolver;

 public:
  Options() {}
  ~Options() {}

  SolverType solver_type{SolverType::GAUSS_NEWTON};   // read by FullBundleAdjustmentSolverRefactor::Solve only
  struct { float threshold_step_size{1e-5}, threshold_cost_change{1e-5}; } convergence_handle;
  struct { float threshold_huber_loss{1.0}, threshold_outlier_rejection{2.0}; } outlier_handle;
  struct { int max_num_iterations{50}; } iteration_handle;
  struct { float initial_lambda{100.0}, decrease_ratio_lambda{0.33f}, increase_ratio_lambda{3.0f}; } trust_region_handle;
  // Extension (not in the reference): false = reference-exact B_ji assignment (last observation of a
  // (pose, point) pair wins, full_bundle_adjustment_solver.cpp:826); true = accumulate (`+=`).
  bool accumulate_offdiagonal_blocks{false};
};

class Summary {
  friend class PoseOnlyBundleAdjustmentSolver;
  friend class FullBundleAdjustmentSolver;
  friend class FullBundleAdjustmentSolverRefactor;

 public:
  Summary() {}
  ~Summary() {}
  std::string BriefReport() {
    const auto default_precision{std::cout.precision()};
    std::stringstream ss;
    ss << "itr " << "  total_cost  " << " avg.reproj. " << " cost_change " << " |step|  " << " |gradient| "
       << " damp_term " << " itr_time[ms] " << "itr_stat\n";
    const size_t num_iterations = optimization_info_list_.size();
    for (size_t iteration = 0; iteration < num_iterations; ++iteration) {
      const OptimizationInfo &info = optimization_info_list_[iteration];
      ss << std::setw(3) << iteration << " ";
      ss << " " << std::scientific << info.cost;
      ss << "    " << std::setprecision(2) << std::scientific << info.average_reprojection_error;
      ss << "    " << std::setprecision(2) << std::scientific << info.cost_change;
      ss << "   " << std::setprecision(2) << std::scientific << info.abs_step;
      ss << "   " << std::setprecision(2) << std::scientific << info.abs_gradient;
      ss << "    " << std::setprecision(2) << std::scientific << info.damping_term;
      ss << "   " << std::setprecision(2) << std::scientific << info.iter_time;
      switch (info.iteration_status) {
        case IterationStatus::UPDATE: ss << "     " << "UPDATE"; break;
        case IterationStatus::SKIPPED: ss << "     " << TEXT_YELLOW(" SKIP "); break;
        case IterationStatus::UPDATE_TRUST_MORE: ss << "     " << TEXT_GREEN("UPDATE"); break;
        default: ss << "     ";
      }
      ss << "\n";
      ss << std::setprecision(default_precision);
    }
    ss << std::setprecision(5);
    ss << "Analytic Solver Report:\n";
    ss << "  Iterations      : " << num_iterations << "\n";
    ss << "  Total time      : " << total_time_in_millisecond_ * 0.001 << " [second]\n";
    if (num_iterations > 0) {  // the reference dereferences front()/back() unguarded (UB when empty)
      ss << "  Initial cost    : " << optimization_info_list_.front().cost << "\n";
      ss << "  Final cost      : " << optimization_info_list_.back().cost << "\n";
      ss << "  Initial reproj. : " << optimization_info_list_.front().average_reprojection_error << " [pixel]\n";
      ss << "  Final reproj.   : " << optimization_info_list_.back().average_reprojection_error << " [pixel]\n";
    }
    ss << ", Termination     : " << (convergence_status_ ? TEXT_GREEN("CONVERGENCE") : TEXT_YELLOW("NO_CONVERGENCE")) << "\n";
    if (max_iteration_ == static_cast<int>(num_iterations))
      ss << TEXT_YELLOW(" WARNIING: MAX ITERATION is reached ! The solution could be local minima.\n");
    ss << std::setprecision(default_precision);
    return ss.str();
  }
  // Declared at solver_option_and_summary.h:83 of the reference and never defined there.  Here: the brief report
  // plus what the brief one leaves out -- the settings the solve ran with, how the trust region behaved (accepted /
  // trusted-more / skipped steps, range of the damping term), where the time went and which test ended the loop.
  std::string FullReport() {
    std::stringstream ss;
    ss << BriefReport();
    const size_t n = optimization_info_list_.size();
    int n_update = 0, n_trust = 0, n_skip = 0;
    double t_sum = 0.0, t_max = 0.0, lam_min = 0.0, lam_max = 0.0;
    for (size_t k = 0; k < n; ++k) {
      const OptimizationInfo &info = optimization_info_list_[k];
      n_update += info.iteration_status == IterationStatus::UPDATE;
      n_trust += info.iteration_status == IterationStatus::UPDATE_TRUST_MORE;
      n_skip += info.iteration_status == IterationStatus::SKIPPED;
      t_sum += info.iter_time;
      t_max = info.iter_time > t_max ? info.iter_time : t_max;
      lam_min = (k == 0 || info.damping_term < lam_min) ? info.damping_term : lam_min;
      lam_max = (k == 0 || info.damping_term > lam_max) ? info.damping_term : lam_max;
    }
    ss << std::setprecision(6);
    ss << "Analytic Solver Full Report:\n";
    ss << "  Settings\n";
    ss << "    max iterations        : " << max_iteration_ << "\n";
    ss << "    threshold step size   : " << threshold_step_size_ << "\n";
    ss << "    threshold cost change : " << threshold_cost_change_ << "\n";
    ss << "  Trust region\n";
    ss << "    steps accepted        : " << (n_update + n_trust) << " (" << n_trust << " with the damping term lowered)\n";
    ss << "    steps skipped         : " << n_skip << "\n";
    ss << "    damping term range    : [" << lam_min << ", " << lam_max << "]\n";
    if (n > 0) {
      const OptimizationInfo &first = optimization_info_list_.front(), &last = optimization_info_list_.back();
      ss << "  Cost\n";
      ss << "    initial -> final      : " << first.cost << " -> " << last.cost;
      if (first.cost != 0.0) ss << "  (x" << last.cost / first.cost << ")";
      ss << "\n";
      ss << "    last cost change      : " << last.cost_change << "\n";
      ss << "    last average step     : " << last.abs_step << "\n";
      ss << "  Time\n";
      ss << "    total                 : " << total_time_in_millisecond_ << " [ms]\n";
      ss << "    iterations (sum, max) : " << t_sum << ", " << t_max << " [ms]\n";
      ss << "    outside the iterations: " << (total_time_in_millisecond_ - t_sum) << " [ms] (finalisation, transfers)\n";
      ss << "  Stopped because         : ";
      if (!convergence_status_) ss << "the iteration limit was reached\n";
      else if (last.abs_step < threshold_step_size_) ss << "the average step size fell below its threshold\n";
      else if (last.cost_change < threshold_cost_change_) ss << "the cost change fell below its threshold\n";
      else ss << "convergence was reported\n";
    }
    return ss.str();
  }
  const double GetTotalTimeInSecond() const { return total_time_in_millisecond_ * 0.001; }
  // read access for tests / callers (the reference exposes these only to its friend solvers)
  const std::vector<OptimizationInfo> &optimization_info_list() const { return optimization_info_list_; }
  bool convergence_status() const { return convergence_status_; }

 protected:   // reference :86-92, filled by the friend solvers
  std::vector<OptimizationInfo> optimization_info_list_;
  int max_iteration_{0};
  double total_time_in_millisecond_{0.0}, threshold_step_size_{0.0}, threshold_cost_change_{0.0};
  bool convergence_status_{true};
};

}  // namespace analytic_solver
}  // namespace visual_navigation
#endif
