/*
 * Headless stand-in for <SDL.h> — TEST INFRASTRUCTURE ONLY.
 *
 * The reference's main.cpp includes <SDL.h> at file scope (main.cpp:4) although none of its
 * hot-path functions (main.cpp:28-139) touch SDL. SDL2 is not installed in this image, so this
 * header declares just the names main.cpp uses (main.cpp:177-381) so that the UNMODIFIED
 * reference translation unit compiles; bodies live in oracle/ref_harness.cpp.
 * Nothing here is shipped in, linked into or called by the product library.
 */
#ifndef ORACLE_STUB_SDL_H
#define ORACLE_STUB_SDL_H
#include <cstdint>
#include <cstdio>

typedef uint8_t  Uint8;
typedef uint32_t Uint32;

struct SDL_Window;
struct SDL_Renderer;
struct SDL_Texture;
struct SDL_PixelFormat { Uint32 Rmask, Gmask, Bmask, Amask; };
struct SDL_Surface { void* pixels; int pitch; SDL_PixelFormat* format; int w, h; };
struct SDL_Rect;

struct SDL_Keysym { int sym; };
struct SDL_KeyboardEvent { SDL_Keysym keysym; };
union SDL_Event { Uint32 type; SDL_KeyboardEvent key; };

enum {
    SDL_INIT_VIDEO = 0x20, SDL_WINDOWPOS_UNDEFINED = 0x1FFF0000, SDL_WINDOW_SHOWN = 4,
    SDL_RENDERER_ACCELERATED = 2, SDL_QUIT = 0x100, SDL_KEYDOWN = 0x300
};
enum {
    SDLK_UP = 1, SDLK_DOWN, SDLK_LEFT, SDLK_RIGHT,
    SDLK_a = 'a', SDLK_s = 's', SDLK_d = 'd', SDLK_w = 'w', SDLK_q = 'q', SDLK_r = 'r'
};

int          SDL_Init(Uint32 flags);
const char*  SDL_GetError();
SDL_Window*  SDL_CreateWindow(const char* title, int x, int y, int w, int h, Uint32 flags);
SDL_Surface* SDL_CreateRGBSurface(Uint32 flags, int w, int h, int depth, Uint32 r, Uint32 g, Uint32 b, Uint32 a);
SDL_Renderer* SDL_CreateRenderer(SDL_Window* w, int index, Uint32 flags);
SDL_Texture* SDL_CreateTextureFromSurface(SDL_Renderer* r, SDL_Surface* s);
void SDL_DestroyWindow(SDL_Window*);
void SDL_DestroyRenderer(SDL_Renderer*);
void SDL_DestroyTexture(SDL_Texture*);
void SDL_FreeSurface(SDL_Surface*);
void SDL_Quit();
int  SDL_PollEvent(SDL_Event* e);
Uint32 SDL_MapRGB(const SDL_PixelFormat* fmt, Uint8 r, Uint8 g, Uint8 b);
Uint32 SDL_MapRGBA(const SDL_PixelFormat* fmt, Uint8 r, Uint8 g, Uint8 b, Uint8 a);
int  SDL_RenderClear(SDL_Renderer*);
int  SDL_RenderCopy(SDL_Renderer*, SDL_Texture*, const SDL_Rect*, const SDL_Rect*);
void SDL_RenderPresent(SDL_Renderer*);

#endif
