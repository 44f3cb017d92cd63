/* ppc_oracle.h -- CPU restatement of the ParticlePhysicsC solver hot path.  TEST INFRASTRUCTURE ONLY;
 * see the header of ppc_oracle.c for scope, pinning and the reference lines each function follows. */
#ifndef PPC_ORACLE_H
#define PPC_ORACLE_H
#include <stddef.h>
#include <stdint.h>

void   ppc_oracle_update( float *px, float *py, float *vx, float *vy, const float *mass, uint8_t *rgba, size_t n, float dt,
                          int do_colour );
void   ppc_oracle_walls( float *px, float *py, float *vx, float *vy, const float *radius, size_t n, uint16_t W, uint16_t H );
void   ppc_oracle_grid_dims( const float *radius, size_t n, uint16_t W, uint16_t H, float *cell_size, int *cols, int *rows );
void   ppc_oracle_grid_build( const float *px, const float *py, size_t n, float cell_size, int cols, int rows, uint32_t *keys,
                              uint32_t *order, uint32_t *cell_start );
size_t ppc_oracle_pairs( const float *px, const float *py, const float *radius, size_t n, int cols, int rows, const uint32_t *keys,
                         const uint32_t *order, const uint32_t *cell_start, int32_t **out );
void   ppc_oracle_free( void *p );
void   ppc_oracle_resolve_sequential( float *px, float *py, float *vx, float *vy, const float *mass, const float *radius, size_t n,
                                      const int32_t *pairs, size_t npairs, float dt );
void   ppc_oracle_resolve_jacobi( float *px, float *py, float *vx, float *vy, const float *mass, const float *radius, size_t n,
                                  const int32_t *pairs, size_t npairs, float dt );
void   ppc_oracle_tilegs_order( const int32_t *pairs, size_t npairs, size_t n, const uint32_t *keys, const uint32_t *cell_start, int cols,
                                int rows, int row0, int wc, int tile_r, int strip_limit, uint32_t *perm_out );
void   ppc_oracle_resolve_permuted( float *px, float *py, float *vx, float *vy, const float *mass, const float *radius, size_t n,
                                    const int32_t *pairs, size_t npairs, float dt, const uint32_t *perm );
void   ppc_oracle_set_tilegs( int wc, int tile_r, int strip_limit );
void   ppc_oracle_collisions_check_grid( float *px, float *py, float *vx, float *vy, const float *mass, const float *radius, size_t n,
                                         uint16_t W, uint16_t H, float dt, int mode, uint32_t *keys_out, uint32_t *order_out,
                                         uint32_t *cell_start_out, int32_t **pairs_out, size_t *npairs_out );
void   ppc_oracle_substep( float *px, float *py, float *vx, float *vy, const float *mass, const float *radius, uint8_t *rgba, size_t n,
                           uint16_t W, uint16_t H, float dt, int mode, int do_colour );
void   ppc_oracle_stats( const float *px, const float *py, const float *vx, const float *vy, const float *mass, const float *radius,
                         size_t n, uint16_t W, uint16_t H, double *out );
#endif
