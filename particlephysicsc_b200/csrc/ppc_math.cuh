// ppc_math.cuh -- per-particle and per-contact arithmetic of the solver, as device functions.
//
// Every expression follows the reference's C evaluation order and operand TYPES, because positions must be
// bit-identical to the reference (one float ulp already exceeds the 1e-4 x radius tolerance in the large boxes,
// SURVEY.md section 7).  The reference is compiled without FMA contraction (-std=c11 => -ffp-contract=off for gcc);
// here that is enforced twice: the library is built with -fmad=false AND parity-critical expressions use the
// explicitly rounded intrinsics (__fmul_rn, __fadd_rn, __dmul_rn, ...), which the compiler never fuses.
// Division and square root are the IEEE-correct forms (__fdiv_rn, __fsqrt_rn); never build with -use_fast_math.
//
// Reference lines are relative to the reference root (src/particles_solver.c unless stated otherwise).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ppc {

// includes/constants.h:10-20 -- the literal TYPES are part of the contract (unsuffixed ones are doubles).
#define PPC_SOLVER_TUNER 0.3f
#define PPC_EPSILON 1e-8f
#define PPC_DEPTH_CAP 0.05f
#define PPC_SLOP_FACTOR 0.01f
#define PPC_BAUMGARTE 0.2f
#define PPC_GRAVITY 9.81f
#define PPC_AIR_RESISTENCE 0.2
#define PPC_RESTITUTION 0.9
#define PPC_DAMPING_WALL 0.9
#define PPC_ENERGY_MIN 1e-4

// ---- particlesUpdate, :20-28 --------------------------------------------------------------------------------
// p += 0.5*a*dt*dt + v*dt in double (the 0.5 literal promotes), v *= (1 - 0.2*dt) in double, v += 9.81f*dt in
// float on BOTH components (gravity on x is reference behaviour, :27).
__device__ __forceinline__ void integrate( float &px, float &py, float &vx, float &vy, const float dt ) {
    const double t  = (double)dt;
    const double ax = __dmul_rn( __dmul_rn( 0.5 * 0, t ), t );                                       // :20
    const double ay = __dmul_rn( __dmul_rn( __dmul_rn( 0.5, (double)PPC_GRAVITY ), t ), t );         // :21
    const double dr = __dsub_rn( 1.0, __dmul_rn( PPC_AIR_RESISTENCE, t ) );                          // :24-25
    const float  g  = __fmul_rn( PPC_GRAVITY, dt );                                                  // :27-28
    px = __double2float_rn( __dadd_rn( (double)px, __dadd_rn( ax, (double)__fmul_rn( vx, dt ) ) ) );
    py = __double2float_rn( __dadd_rn( (double)py, __dadd_rn( ay, (double)__fmul_rn( vy, dt ) ) ) );
    vx = __double2float_rn( __dmul_rn( (double)vx, dr ) );
    vy = __double2float_rn( __dmul_rn( (double)vy, dr ) );
    vx = __fadd_rn( vx, g );
    vy = __fadd_rn( vy, g );
}

// ---- colour, :31-39 with src/utils.c:7-45 -------------------------------------------------------------------
__device__ __forceinline__ unsigned char byte_of( const float intensity ) {   // src/utils.c:7-16
    if ( intensity < 0.0f )
        return 0;
    if ( intensity > 255.0f )
        return 255;
    return (unsigned char)( __float2int_rz( intensity ) & 0xff );
}

__device__ __forceinline__ uchar4 energy_colour( const float vx, const float vy, const float mass ) {
    const float speed = __fadd_rn( __fmul_rn( vx, vx ), __fmul_rn( vy, vy ) );                                   // :31
    float t = __double2float_rn( __ddiv_rn( __dadd_rn( (double)__fmul_rn( mass, speed ), PPC_ENERGY_MIN ),
                                            ( 1000000 - PPC_ENERGY_MIN ) ) );                                     // :32
    if ( t < 0.0f )                                                                                              // :34
        t = 0.0f;
    else if ( t > 1.0f )
        t = 1.0f;
    t = __double2float_rn( pow( (double)t, (double)( 1.0f / 3.0f ) ) );                                          // :35-36
    const float hue   = __fmul_rn( 240.0f, __fsub_rn( 1.0f, t ) );                                               // :38
    const float arg   = __fdiv_rn( __fmul_rn( hue, 3.14159265f ), 180.0f );                                      // utils.c:31
    const float c     = __double2float_rn( cos( (double)arg ) );
    const float s     = __double2float_rn( sin( (double)arg ) );
    const float r13   = __fsqrt_rn( 1.0f / 3.0f );
    const float omc   = __fsub_rn( 1.0f, c );
    const float third = 1.0f / 3.0f;
    const float diag  = __fadd_rn( c, __fmul_rn( omc, third ) );                                                  // utils.c:37-39
    const float offm  = __fsub_rn( __fmul_rn( third, omc ), __fmul_rn( r13, s ) );
    const float offp  = __fadd_rn( __fmul_rn( third, omc ), __fmul_rn( r13, s ) );
    const float R = 230.0f, G = 41.0f, B = 55.0f;   // Raylib RED
    uchar4 out;
    out.x = byte_of( __fadd_rn( __fadd_rn( __fmul_rn( R, diag ), __fmul_rn( G, offm ) ), __fmul_rn( B, offp ) ) );
    out.y = byte_of( __fadd_rn( __fadd_rn( __fmul_rn( R, offp ), __fmul_rn( G, diag ) ), __fmul_rn( B, offm ) ) );
    out.z = byte_of( __fadd_rn( __fadd_rn( __fmul_rn( R, offm ), __fmul_rn( G, offp ) ), __fmul_rn( B, diag ) ) );
    out.w = 255;                                                                                                  // :39
    return out;
}

// ---- wall pass, :207-228 ------------------------------------------------------------------------------------
// All four tests use the position loaded at the top; W and H arrive as (float)(int)uint16.
__device__ __forceinline__ void walls( float &px, float &py, float &vx, float &vy, const float r, const float W, const float H ) {
    const float x = px, y = py;
    const float xr = __fsub_rn( W, r ), yb = __fsub_rn( H, r );
    if ( x <= r ) {
        px = r;
        vx = __double2float_rn( __dmul_rn( (double)vx, -PPC_DAMPING_WALL ) );
    }
    if ( x >= xr ) {
        px = xr;
        vx = __double2float_rn( __dmul_rn( (double)vx, -PPC_DAMPING_WALL ) );
    }
    if ( y <= r ) {
        py = r;
        vy = __double2float_rn( __dmul_rn( (double)vy, -PPC_DAMPING_WALL ) );
    }
    if ( y >= yb ) {
        py = yb;
        vy = __double2float_rn( __dmul_rn( (double)vy, -PPC_DAMPING_WALL ) );
    }
}

// ---- cell coordinate, :280-289 ------------------------------------------------------------------------------
// One IEEE division (a reciprocal multiply flips keys), truncation toward zero, clamp.  NaN / out-of-range
// follow x86's cvttss2si (INT_MIN -> cell 0), which is what the compiled reference does.
__device__ __forceinline__ int cell_coord( const float p, const float cell_size, const int limit ) {
    const float q = __fdiv_rn( p, cell_size );
    int         c = ( q >= -2147483648.0f && q < 2147483648.0f ) ? __float2int_rz( q ) : INT32_MIN;
    if ( c < 0 )
        c = 0;
    if ( c >= limit )
        c = limit - 1;
    return c;
}

// ---- pair test, :352-362 ------------------------------------------------------------------------------------
// d = p_i - p_j.  The test is symmetric under swapping i and j (negation and the float sum of radii are exact /
// commutative), so either particle of a pair may evaluate it.
__device__ __forceinline__ bool overlaps( const float dx, const float dy, const float rsum ) {
    if ( dx > rsum || dx < -rsum )
        return false;
    if ( dy > rsum || dy < -rsum )
        return false;
    const float d2 = __fadd_rn( __fmul_rn( dx, dx ), __fmul_rn( dy, dy ) );
    return d2 <= __fmul_rn( rsum, rsum );
}

// ---- contact response of ONE pair (lo = lower API index = the reference's i, hi = its j) ----------------------
// Pass 1 (:84-127): geometry frozen before anything moves.
struct ContactGeom {
    float nx, ny;      // unit normal from hi towards lo (:96-97)
    float raw;         // penetration depth (:103)
    float corrected;   // clamped share resolved positionally (:106-113)
    float vbias;       // Baumgarte velocity term (:116-127)
};

__device__ __forceinline__ ContactGeom contact_geom( const float xlo, const float ylo, const float rlo, const float xhi, const float yhi,
                                                     const float rhi, const float dt ) {
    ContactGeom C;
    const float dx = __fsub_rn( xlo, xhi ), dy = __fsub_rn( ylo, yhi );                      // :85-86
    float       d2 = __fadd_rn( __fmul_rn( dx, dx ), __fmul_rn( dy, dy ) );                  // :87
    d2             = fmaxf( d2, PPC_EPSILON );                                               // :88
    const float d  = __fsqrt_rn( d2 );                                                       // :89
    if ( d < PPC_EPSILON ) {                                                                 // :92 (unreachable: d >= 1e-4)
        C.nx = 1.0f;
        C.ny = 0.0f;
    } else {
        C.nx = __fdiv_rn( dx, d );                                                           // :96-97
        C.ny = __fdiv_rn( dy, d );
    }
    C.raw       = __fsub_rn( __fadd_rn( rlo, rhi ), d );                                     // :103
    C.corrected = 0.0f;
    C.vbias     = 0.0f;
    if ( C.raw > 0.0f ) {
        const float cap = __fmul_rn( PPC_DEPTH_CAP, fminf( rlo, rhi ) );                     // :107
        float       cor = __fmul_rn( C.raw, PPC_SOLVER_TUNER );                              // :108
        cor             = fmaxf( -cap, fminf( cor, cap ) );                                  // :109
        C.corrected     = cor;
        const float residual = __fsub_rn( C.raw, cor );                                      // :117
        const float slop     = __fmul_rn( PPC_SLOP_FACTOR, fmaxf( rlo, rhi ) );              // :118
        if ( residual > slop )
            C.vbias = __fdiv_rn( __fmul_rn( PPC_BAUMGARTE, __fsub_rn( residual, slop ) ), dt );   // :121
    }
    return C;
}

__device__ __forceinline__ float inv_mass_u8( const float m ) {   // :141-142,146-147 masses are read through uint8_t
    return __fdiv_rn( 1.0f, (float)( (unsigned)__float2int_rz( m ) & 0xffu ) );
}

// Pass 2, positional part (:145-156): lo gets +(dpx_lo, dpy_lo), hi gets -(dpx_hi, dpy_hi); only when push.
struct PushTerms {
    float dpx_lo, dpy_lo, dpx_hi, dpy_hi;
    bool  push;
};
__device__ __forceinline__ PushTerms push_terms( const ContactGeom &C, const float wlo, const float whi ) {
    PushTerms T;
    T.push = C.raw > 0.0f;                                                                   // :145
    T.dpx_lo = T.dpy_lo = T.dpx_hi = T.dpy_hi = 0.0f;
    if ( T.push ) {
        const float wsum = __fadd_rn( wlo, whi );                                            // :148
        const float flo  = __fmul_rn( __fdiv_rn( wlo, wsum ), C.corrected );                 // :149-150
        const float fhi  = __fmul_rn( __fdiv_rn( whi, wsum ), C.corrected );
        T.dpx_lo = __fmul_rn( flo, C.nx );                                                   // :152-155
        T.dpy_lo = __fmul_rn( flo, C.ny );
        T.dpx_hi = __fmul_rn( fhi, C.nx );
        T.dpy_hi = __fmul_rn( fhi, C.ny );
    }
    return T;
}

// Pass 2, velocity part (:158-190) from the velocities handed in: lo gets +(dvx_lo, dvy_lo), hi gets -(dvx_hi, dvy_hi);
// only when impulse (the :170 early-out did not fire).
struct ImpulseTerms {
    float dvx_lo, dvy_lo, dvx_hi, dvy_hi;
    bool  impulse;
};
__device__ __forceinline__ ImpulseTerms impulse_terms( const ContactGeom &C, const float wlo, const float whi, const float vxlo,
                                                       const float vylo, const float vxhi, const float vyhi ) {
    ImpulseTerms T;
    const float  rx = __fsub_rn( vxlo, vxhi ), ry = __fsub_rn( vylo, vyhi );                 // :163-164
    const float  vn = __fadd_rn( __fmul_rn( rx, C.nx ), __fmul_rn( ry, C.ny ) );             // :166
    T.impulse = !( vn > 0.0f && C.vbias == 0.0f );                                           // :170
    T.dvx_lo = T.dvy_lo = T.dvx_hi = T.dvy_hi = 0.0f;
    if ( T.impulse ) {
        const float want = __double2float_rn( __dadd_rn( __dmul_rn( -PPC_RESTITUTION, (double)vn ), (double)C.vbias ) );   // :175
        const float wsum = __fadd_rn( wlo, whi );                                            // :179
        const float jmag = __fdiv_rn( __fsub_rn( want, vn ), wsum );                         // :180
        const float ix = __fmul_rn( jmag, C.nx ), iy = __fmul_rn( jmag, C.ny );              // :183-184
        T.dvx_lo = __fmul_rn( wlo, ix );                                                     // :187-190
        T.dvy_lo = __fmul_rn( wlo, iy );
        T.dvx_hi = __fmul_rn( whi, ix );
        T.dvy_hi = __fmul_rn( whi, iy );
    }
    return T;
}

}   // namespace ppc
