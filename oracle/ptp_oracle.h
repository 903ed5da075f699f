/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference's PIC step
 * (PenningTrap::movePlasmas and everything below it). See ptp_oracle.c. */
#ifndef PTP_ORACLE_H
#define PTP_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ptpo_trap ptpo_trap;

ptpo_trap* ptpo_trap_create(double radius, int nElectrodes, const double* lengths, const double* potentials,
	int nGaps, const double* gaps, int Nz, int Nr);
void ptpo_trap_destroy(ptpo_trap* t);
void ptpo_trap_info(const ptpo_trap* t, int* Nz, int* Nr, double* hz, double* hr, double* length, double* radius);
void ptpo_set_electrode(ptpo_trap* t, int index, double potential);
void ptpo_matrix_apply(const ptpo_trap* t, const double* x, double* y);
void ptpo_wall_potential(const ptpo_trap* t, double* vWall);
void ptpo_wall_rhs(const ptpo_trap* t, double* rhs);
void ptpo_solve(const ptpo_trap* t, const double* rhs, double* phi);
void ptpo_well_limits(const ptpo_trap* t, const double* phiTrap, int* left, int* right);
void ptpo_node_efield(const ptpo_trap* t, int nPhi, const double* const* phis, double* eNodes);
double ptpo_gather(const ptpo_trap* t, const double* eNodes, int r, double z);
long ptpo_move_rings(const ptpo_trap* t, const double* eNodes, long n, int* r, double* z, double* v,
	double dt, double charge, double mass);
void ptpo_deposit(const ptpo_trap* t, long n, const int* r, const double* z, double macroChargeDensity, double* rhs);
void ptpo_cell_index(const ptpo_trap* t, long n, const int* r, const double* z, int* k, int* idx);
double ptpo_potential_energy(const ptpo_trap* t, int nPhi, const double* const* phis, long n, const int* r,
	const double* z, double chargeMacro);

#ifdef __cplusplus
}
#endif
#endif
