/*
 * Minimal stand-in for <raylib.h>, test infrastructure only.
 *
 * The reference solver (src/particles_solver.c, src/utils.c, src/particles_arrays.c) touches Raylib
 * only through the Color / Vector2 / Texture2D / Image value types, four colour literals and five
 * image/texture/draw calls (src/particles_arrays.c:297-307,337).  Raylib itself is absent from this
 * image, so the oracle build (oracle/Makefile) compiles the UNMODIFIED reference sources against this
 * header.  Types and literal values follow raylib's public header (raylib.h, v4/v5: Color is 4 x u8,
 * Vector2 is 2 x float, RED = {230,41,55,255}, BLUE = {0,121,241,255}, RAYWHITE = {245,245,245,255}).
 * The draw calls are declared but only defined as aborting stubs in oracle/ref_shim.c.
 */
#ifndef RAYLIB_H
#define RAYLIB_H

typedef struct Color {
    unsigned char r, g, b, a;
} Color;

typedef struct Vector2 {
    float x, y;
} Vector2;

typedef struct Image {
    void *data;
    int   width, height, mipmaps, format;
} Image;

typedef struct Texture {
    unsigned int id;
    int          width, height, mipmaps, format;
} Texture;
typedef Texture Texture2D;

#define WHITE    ( (Color){ 255, 255, 255, 255 } )
#define BLACK    ( (Color){ 0, 0, 0, 255 } )
#define RED      ( (Color){ 230, 41, 55, 255 } )
#define BLUE     ( (Color){ 0, 121, 241, 255 } )
#define RAYWHITE ( (Color){ 245, 245, 245, 255 } )

Image     GenImageColor( int width, int height, Color color );
void      ImageDrawCircle( Image *dst, int centerX, int centerY, int radius, Color color );
Texture2D LoadTextureFromImage( Image image );
void      UnloadImage( Image image );
void      DrawTextureEx( Texture2D texture, Vector2 position, float rotation, float scale, Color tint );

#endif
