// Driver with the reference's command line: HiFiLES <input_file>  (reference src/HiFiLES.cpp:41-343).
// Output is reduced to what the hot path's parity comparators need: the residual table on stdout and history.plt
// (reference src/output.cpp:2250-2408).  Plot / restart / probe writers are out of scope.
#include "hifiles.h"
#include <cstdio>
#include <ctime>

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
    FILE *hist = fopen("history.plt", "w");
    clock_t init_time = clock();
    int i_steps = 0;
    while (i_steps < run_input.n_steps)
    {
      calc_time_step(&FlowSol);
      for (int i = 0; i < RKSteps; i++)
      {
        CalcResidual(FlowSol.ini_iter + i_steps, i, &FlowSol);
        for (int j = 0; j < FlowSol.n_ele_types; j++) FlowSol.mesh_eles(j)->AdvanceSolution(i, run_input.adv_type);
        if (run_input.shock_cap)
          for (int j = 0; j < FlowSol.n_ele_types; j++) FlowSol.mesh_eles(j)->shock_capture();
      }
      FlowSol.time += run_input.dt;
      run_input.time = FlowSol.time;
      i_steps++;
      if (i_steps == 1 || i_steps % run_input.monitor_res_freq == 0)
      {
        CalcNormResidual(&FlowSol);
        if (i_steps == 1) printf("\n  Iter       Res[Rho]   Res[RhoVelx]   Res[RhoVely]%s      Res[RhoE]\n", FlowSol.n_dims == 3 ? "   Res[RhoVelz]" : "");
        printf("%6d", FlowSol.ini_iter + i_steps);
        for (int f = 0; f < n_fields; f++) printf(" %14.8f", FlowSol.norm_residual(f));
        printf("\n");
        if (hist)
        {
          fprintf(hist, "%d", FlowSol.ini_iter + i_steps);
          for (int f = 0; f < n_fields; f++) fprintf(hist, ", %.15g", log10(FlowSol.norm_residual(f)));
          fprintf(hist, ", %.15g\n", (double)(clock() - init_time) / CLOCKS_PER_SEC / 60.);
        }
      }
    }
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
