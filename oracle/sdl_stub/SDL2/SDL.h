/* Headless stand-in for <SDL2/SDL.h>: just enough declarations for the
 * reference's unmodified src/main.cpp (usage sites main.cpp:10,91-93,302-363,
 * 558-628) to compile where SDL2 is not installed. SDL_Init fails, so the
 * interactive viewer refuses to start; the --ray-file benchmark mode never
 * touches SDL. Test/bench infrastructure only. */
#ifndef HGB_SDL_STUB_H
#define HGB_SDL_STUB_H
#include <stdint.h>
typedef struct SDL_Surface { int w, h, pitch; void* pixels; } SDL_Surface;
typedef struct SDL_Window SDL_Window;
typedef int SDL_bool;
enum { SDL_FALSE = 0, SDL_TRUE = 1 };
enum { SDL_FIRSTEVENT = 0, SDL_QUIT = 0x100, SDL_KEYDOWN = 0x300, SDL_KEYUP, SDL_MOUSEMOTION = 0x400,
       SDL_MOUSEBUTTONDOWN, SDL_MOUSEBUTTONUP, SDL_LASTEVENT = 0xFFFF };
enum { SDLK_ESCAPE = 27, SDLK_c = 'c', SDLK_m = 'm', SDLK_RIGHT = 0x4000004F, SDLK_LEFT, SDLK_DOWN, SDLK_UP,
       SDLK_KP_MINUS = 0x40000056, SDLK_KP_PLUS };
#define SDL_INIT_VIDEO 0x20u
#define SDL_WINDOWPOS_UNDEFINED 0x1FFF0000
typedef struct SDL_Keysym { int sym; } SDL_Keysym;
typedef struct SDL_KeyboardEvent { uint32_t type; SDL_Keysym keysym; } SDL_KeyboardEvent;
typedef struct SDL_MouseMotionEvent { uint32_t type; int xrel, yrel; } SDL_MouseMotionEvent;
typedef union SDL_Event { uint32_t type; SDL_KeyboardEvent key; SDL_MouseMotionEvent motion; } SDL_Event;
static inline int SDL_Init(uint32_t) { return -1; }
static inline void SDL_Quit(void) {}
static inline SDL_Window* SDL_CreateWindow(const char*, int, int, int, int, uint32_t) { return 0; }
static inline void SDL_DestroyWindow(SDL_Window*) {}
static inline SDL_Surface* SDL_GetWindowSurface(SDL_Window*) { return 0; }
static inline int SDL_UpdateWindowSurface(SDL_Window*) { return 0; }
static inline void SDL_SetWindowTitle(SDL_Window*, const char*) {}
static inline int SDL_LockSurface(SDL_Surface*) { return 0; }
static inline void SDL_UnlockSurface(SDL_Surface*) {}
static inline void SDL_FlushEvents(uint32_t, uint32_t) {}
static inline int SDL_PollEvent(SDL_Event*) { return 0; }
static inline int SDL_SetRelativeMouseMode(SDL_bool) { return 0; }
static inline uint32_t SDL_GetTicks(void) { return 0; }
#endif
