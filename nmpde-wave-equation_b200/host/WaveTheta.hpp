// WaveTheta.hpp -- theta-method integrator with the reference's constructor and run()
// (include/WaveTheta.hpp:46-194).  All numerics run on the GPU through libwavegpu.
#ifndef WAVE_THETA_HPP
#define WAVE_THETA_HPP

#include "WaveEquationBase.hpp"

class WaveTheta : public WaveEquationBase
{
  public:
    WaveTheta(const std::string& problem_name_,
              const std::pair<unsigned int, unsigned int>& N_el_,
              const std::pair<Point<dim>, Point<dim>>& geometry_,
              const unsigned int& r_,
              const double& T_,
              const double& theta_,
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
          theta(theta_)
    {
    }

    void run() override;

  protected:
    void setup();
    void assemble_matrices();

    const double theta;
    unsigned int current_iterations_u = 0;
    unsigned int current_iterations_v = 0;
};

#endif
