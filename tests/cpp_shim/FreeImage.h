// Test-only stand-in for <FreeImage.h> (not installed here): just the declarations the reference's FreeImageHelper.h / VirtualSensor.h
// need so that main.cpp and experiment.cpp can be COMPILED against the drop-in headers (tests/test_reference_drivers_compile.py).
// Nothing is implemented: the object files are never linked or run.
#pragma once
typedef unsigned char BYTE;
struct FIBITMAP;
struct RGBQUAD { BYTE rgbBlue, rgbGreen, rgbRed, rgbReserved; };
enum FREE_IMAGE_FORMAT { FIF_UNKNOWN = -1, FIF_PNG = 13 };
enum FREE_IMAGE_FILTER { FILTER_CATMULLROM = 4 };
void FreeImage_Initialise(int load_local_plugins_only = 0);
FREE_IMAGE_FORMAT FreeImage_GetFileType(const char*, int);
FREE_IMAGE_FORMAT FreeImage_GetFIFFromFilename(const char*);
int FreeImage_FIFSupportsReading(FREE_IMAGE_FORMAT);
FIBITMAP* FreeImage_Load(FREE_IMAGE_FORMAT, const char*, int flags = 0);
FIBITMAP* FreeImage_ConvertToRGBAF(FIBITMAP*);
FIBITMAP* FreeImage_Rescale(FIBITMAP*, int, int, FREE_IMAGE_FILTER);
FIBITMAP* FreeImage_Allocate(int, int, int);
void FreeImage_Unload(FIBITMAP*);
unsigned FreeImage_GetWidth(FIBITMAP*);
unsigned FreeImage_GetHeight(FIBITMAP*);
BYTE* FreeImage_GetBits(FIBITMAP*);
int FreeImage_SetPixelColor(FIBITMAP*, unsigned, unsigned, RGBQUAD*);
int FreeImage_Save(FREE_IMAGE_FORMAT, FIBITMAP*, const char*, int);
