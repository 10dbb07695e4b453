// Driver with the reference's command line: HiFiLES <input_file>  (reference src/HiFiLES.cpp:41-343).
// Output is reduced to what the hot path's parity comparators need: the residual table on stdout and history.plt
// (reference src/output.cpp:2250-2408).  Paraview files (plot.cpp) and ASCII restart files are written as the reference does; Tecplot / CGNS / probe writers are not built.
#include <unistd.h>
#include "hifiles.h"
#include <cstdio>
#include <ctime>
#include <fstream>

using namespace std;

int main(int argc, char *argv[])
{
  if (argc < 2)
  {
    cout << "No input file specified. For help use -h or --help " << endl;
    return 0;
  }
  if (!strcmp(argv[1], "-h") || !strcmp(argv[1], "-help"))
  {
    cout << "To run, use HiFiLES <input_file>" << endl;
    return 0;
  }
  try
  {
    struct solution FlowSol;
    mesh mesh_data;
    run_input.setup(argv[1], 0);
    SetInput(&FlowSol);
    GeoPreprocess(&FlowSol, mesh_data);
    InitSolution(&FlowSol);
    int RKSteps = get_n_rk_steps(run_input.adv_type);
    int n_fields = (run_input.equation == 0) ? FlowSol.n_dims + 2 : 1;
    // a restarted run appends to an existing history file (reference src/output.cpp:2277-2288)
    const bool append = run_input.restart_flag != 0 && access("history.plt", W_OK) != -1;
    FILE *hist = fopen("history.plt", append ? "a" : "w");
    const int n_diags = run_input.n_integral_quantities;
    if (hist && !append)
    {
      // header of output::HistoryOutput (reference src/output.cpp:2299-2343)
      fprintf(hist, "TITLE = \"HiFiLES simulation\"\nVARIABLES = \"Iteration\"");
      if (run_input.equation == 0)
      {
        if (FlowSol.n_dims == 2)
          fprintf(hist, ",\"log<sub>10</sub>(Res[<greek>r</greek>])\",\"log<sub>10</sub>(Res[<greek>r</greek>v<sub>x</sub>])\",\"log<sub>10</sub>(Res[<greek>r</greek>v<sub>y</sub>])\",\"log<sub>10</sub>(Res[<greek>r</greek>E])\"");
        else
          fprintf(hist, ",\"log<sub>10</sub>(Res[<greek>r</greek>])\",\"log<sub>10</sub>(Res[<greek>r</greek>v<sub>x</sub>])\",\"log<sub>10</sub>(Res[<greek>r</greek>v<sub>y</sub>])\",\"log<sub>10</sub>(Res[<greek>r</greek>v<sub>z</sub>])\",\"log<sub>10</sub>(Res[<greek>r</greek>E])\"");
      }
      else
        fprintf(hist, ",\"log<sub>10</sub>(Res[<greek>r</greek>])\"");
      if (run_input.equation == 0 && run_input.calc_force)
        fprintf(hist, FlowSol.n_dims == 2 ? ",\"F<sub>x</sub>(Total)\",\"F<sub>y</sub>(Total)\",\"CL</sub>(Total)\",\"CD</sub>(Total)\""
                                          : ",\"F<sub>x</sub>(Total)\",\"F<sub>y</sub>(Total)\",\"F<sub>z</sub>(Total)\",\"CL</sub>(Total)\",\"CD</sub>(Total)\"");
      for (int i = 0; i < n_diags; i++) fprintf(hist, ",\"Diagnostics[%s]\"", run_input.integral_quantities(i).c_str());
      fprintf(hist, ",\"Time<sub>Physical</sub>(sec)\",\"Time<sub>Comp</sub>(m)\"\nZONE T= \"Convergence history\"\n");
    }
    clock_t init_time = clock();
    int i_steps = 0;
    /*! Dump initial Paraview or Tecplot file (reference src/HiFiLES.cpp:171-182; the CGNS writer is not built) */
    write_plot(FlowSol.ini_iter + i_steps, &FlowSol);
    while (i_steps < run_input.n_steps)
    {
      calc_time_step(&FlowSol);
      if (run_input.pressure_ramp) upload_bc_table(&FlowSol);
      // the residual (and the gradient behind the integral diagnostics) of the last stage is read after monitored steps
      const bool monitored = (i_steps + 1 == 1) || ((i_steps + 1) % run_input.monitor_res_freq == 0) ||
                             (run_input.n_diagnostic_fields > 0 && (i_steps + 1) % run_input.plot_freq == 0);
      for (int i = 0; i < RKSteps; i++) AdvanceStage(FlowSol.ini_iter + i_steps, i, &FlowSol, monitored && i == RKSteps - 1);
      FlowSol.time += run_input.dt;
      run_input.time = FlowSol.time;
      i_steps++;
      if (run_input.pressure_ramp) run_input.ramp_counter++;
      /*! Compute time-averaged quantities (reference src/HiFiLES.cpp:241-245) */
      if (i_steps == 1) run_input.spinup_time = FlowSol.time; // set start time for averaging
      if (run_input.n_average_fields) CalcTimeAverageQuantities(&FlowSol);
      if (i_steps == 1 || i_steps % run_input.monitor_res_freq == 0)
      {
        /*! Compute the value of the forces (reference src/HiFiLES.cpp:250-254) */
        if (run_input.calc_force != 0) CalcForces(FlowSol.ini_iter + i_steps, (i_steps == 1 || i_steps % run_input.monitor_cp_freq == 0), &FlowSol);
        if (n_diags) CalcIntegralQuantities(&FlowSol);
        CalcNormResidual(&FlowSol);
        if (i_steps == 1)
          printf("\n  Iter       Res[Rho]   Res[RhoVelx]   Res[RhoVely]%s      Res[RhoE]%s\n", FlowSol.n_dims == 3 ? "   Res[RhoVelz]" : "",
                 !run_input.calc_force ? "" : (FlowSol.n_dims == 3 ? "       Fx_Total       Fy_Total       Fz_Total" : "       Fx_Total       Fy_Total"));
        printf("%6d", FlowSol.ini_iter + i_steps);
        for (int f = 0; f < n_fields; f++) printf(" %14.8f", FlowSol.norm_residual(f));
        if (run_input.calc_force != 0)
          for (int d = 0; d < FlowSol.n_dims; d++) printf(" %14.8f", FlowSol.inv_force(d) + FlowSol.vis_force(d));
        printf("\n");
        if (hist)
        {
          fprintf(hist, "%d", FlowSol.ini_iter + i_steps);
          for (int f = 0; f < n_fields; f++) fprintf(hist, ", %.15g", log10(FlowSol.norm_residual(f)));
          if (run_input.calc_force != 0)
          {
            for (int d = 0; d < FlowSol.n_dims; d++) fprintf(hist, ", %.15g", FlowSol.inv_force(d) + FlowSol.vis_force(d));
            fprintf(hist, ", %.15g, %.15g", FlowSol.coeff_lift, FlowSol.coeff_drag);
          }
          for (int q = 0; q < n_diags; q++) fprintf(hist, ", %.15g", FlowSol.integral_quantities(q));
          fprintf(hist, ", %.15g", (run_input.viscous && run_input.equation == 0) ? FlowSol.time * run_input.time_ref : FlowSol.time);
          fprintf(hist, ", %.15g\n", (double)(clock() - init_time) / CLOCKS_PER_SEC / 60.);
        }
      }
      if (i_steps % run_input.plot_freq == 0) write_plot(FlowSol.ini_iter + i_steps, &FlowSol);
      if (i_steps % run_input.restart_dump_freq == 0) write_restart_ascii(&FlowSol, FlowSol.ini_iter + i_steps);
    }
    /*! Calculate Error (reference src/HiFiLES.cpp:324-325) */
    if (run_input.test_case) compute_error(FlowSol.ini_iter + i_steps, &FlowSol);
    if (hist) fclose(hist);
    hf_check(hf_dev_sync(FlowSol.ctx));
    printf("Execution time= %f s\n", (double)(clock() - init_time) / CLOCKS_PER_SEC);
  }
  catch (const std::exception &e)
  {
    cout << e.what() << endl;
    return 1;
  }
  return 0;
}
