/* Entry stub (TEST INFRASTRUCTURE): runs the reference's own main(), which the
 * oracle build renames to citcom_main (Citcom.c:54), so `_ref/citcom_ref input`
 * behaves like the reference's citcom.mpi on top of the MPI shim. */
int citcom_main(int argc, char **argv);
int main(int argc, char **argv) { return citcom_main(argc, argv); }
