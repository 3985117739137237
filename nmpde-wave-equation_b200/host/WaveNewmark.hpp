// WaveNewmark.hpp -- Newmark-beta integrator with the reference's constructor and run()
// (include/WaveNewmark.hpp:38-180).  All numerics run on the GPU through libwavegpu.
#ifndef WAVE_NEWMARK_HPP
#define WAVE_NEWMARK_HPP

#include "WaveEquationBase.hpp"

class WaveNewmark : public WaveEquationBase
{
  public:
    WaveNewmark(const std::string& problem_name_,
                const std::pair<unsigned int, unsigned int>& N_el_,
                const std::pair<Point<dim>, Point<dim>>& geometry_,
                const unsigned int& r_,
                const double& T_,
                const double& gamma_,
                const double& beta_,
                const double& delta_t_,
                const Function<dim>& c_,
                Function<dim>& f_,
                const Function<dim>& u0_,
                const Function<dim>& v0_,
                Function<dim>& g_,
                Function<dim>& dgdt_,
                const unsigned int log_every_ = 10,
                const unsigned int print_every_ = 10,
                Function<dim>* exact_solution_ = nullptr)
        : WaveEquationBase(problem_name_, N_el_, geometry_, r_, T_, delta_t_, c_, f_, u0_, v0_, g_, dgdt_,
                           log_every_, print_every_, exact_solution_),
          gamma(gamma_), beta(beta_)
    {
    }

    void run() override;

  protected:
    void setup();
    void assemble_matrices();

    const double gamma;
    const double beta;
    unsigned int current_iterations = 0;
};

#endif
