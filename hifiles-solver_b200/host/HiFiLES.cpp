// Driver with the reference's command line: HiFiLES <input_file>  (reference src/HiFiLES.cpp:41-343).
// Output is reduced to what the hot path's parity comparators need: the residual table on stdout and history.plt
// (reference src/output.cpp:2250-2408).  Paraview files (plot.cpp) and ASCII restart files are written as the reference does; Tecplot / CGNS / probe writers are not built.
#include <unistd.h>
#include "hifiles.h"
#include <cstdio>
#include <ctime>
#include <fstream>

using namespace std;

// ASCII restart file Rest_<iter>_p0000.dat: time, then per element type a header with the solution-point set and the
// solution of every element under its global id, 15 significant digits (reference src/output.cpp:1753-1818,
// src/eles.cpp:845-869, <type>::write_restart_info_ascii).  It is one of the reference's parity comparators.
static void write_restart_ascii(struct solution *FlowSol, int in_file_num)
{
  char name[256];
  if (FlowSol->nproc > 1) FatalError("restart files of partitioned runs are outside the scope of this build");
  snprintf(name, sizeof(name), "Rest_%.09d_p%.04d.dat", in_file_num, 0);
  cout << "Writing Restart file for step " << in_file_num << " ...." << flush;
  ofstream f(name);
  f.precision(15);
  f << FlowSol->time << endl;
  for (int t = 0; t < FlowSol->n_ele_types; t++)
  {
    eles *e = FlowSol->mesh_eles(t);
    if (e->get_n_eles() == 0) continue;
    e->cp_disu_upts_gpu_cpu();
    const int type = e->get_ele_type();
    static const char *title[5] = {"TRIS", "QUADS", "TETS", "PRIS", "HEXAS"};
    static const char *count[5] = {"Number of solution points per triangular element", "Number of solution points per quadrilateral element",
                                   "Number of solution points per element", "Number of solution points per prismatic element",
                                   "Number of solution points per hexahedral element"};
    f << title[type] << endl << "Order" << endl << e->order << endl << count[type] << endl << e->n_upts_per_ele << endl;
    if (type == QUAD || type == HEX)
    {
      f << "Location of solution points in 1D" << endl;
      for (int i = 0; i < e->order + 1; ++i) f << e->loc_1d_upts(i) << " ";
      f << endl;
    }
    else if (type == PRISM)
    {
      eles_pris *p = static_cast<eles_pris *>(e);
      f << "Number of solution points in triangle" << endl << p->n_upts_tri << endl;
      f << "Location of solution points in 1D" << endl;
      for (int i = 0; i < e->order + 1; ++i) f << p->loc_upts_pri_1d(i) << " ";
      f << endl;
      f << "Location of solution points in triangle" << endl;
      for (int i = 0; i < p->n_upts_tri; i++)
      {
        for (int j = 0; j < 2; j++) f << p->loc_upts_pri_tri(j, i) << " ";
        f << endl;
      }
    }
    else
    {
      f << (type == TRI ? "Location of solution points in triangular elements" : "Location of solution points in tetrahedral elements") << endl;
      for (int i = 0; i < e->n_upts_per_ele; i++)
      {
        for (int j = 0; j < e->n_dims; j++) f << e->loc_upts(j, i) << " ";
        f << endl;
      }
    }
    f << "n_eles" << endl << e->n_eles << endl << "ele2global_ele hf_array" << endl;
    for (int i = 0; i < e->n_eles; i++) f << e->ele2global_ele(i) << " ";
    f << endl << "data" << endl;
    for (int i = 0; i < e->n_eles; i++)
    {
      f << e->ele2global_ele(i) << endl;
      for (int j = 0; j < e->n_upts_per_ele; j++)
      {
        for (int k = 0; k < e->n_fields; k++) f << e->disu_upts(0)(j, i, k) << " ";
        f << endl;
      }
    }
    f << endl;
  }
  cout << "done" << endl;
}

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
