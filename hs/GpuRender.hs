{-# LANGUAGE ForeignFunctionInterface #-}
{-# LANGUAGE NamedFieldPuns #-}
{- |
GpuRender.hs -- the reference-side binding of libyahr_b200.so.

This is the module a yahr maintainer would add next to main.hs to get a "gpu" parallel mode.
It is SOURCE ONLY: no GHC exists in the build image, so it has never been compiled there; it is
kept small and uses only base + vector (already dependencies of yahr.cabal:22-24).

  main.hs:134-137 gains one case:      "gpu" -> samplesFromGpu
  and the image is built from the frame directly (JuicyPixels' `Image PixelRGBF` is
  {width, height, Storable vector of RGB floats, row-major, top row first} -- exactly the layout
  yahr_b200_render writes), skipping samplesToImage.

Build:  add `extra-libraries: yahr_b200`, `extra-lib-dirs: <repo>/yahr_b200` and
        `other-modules: GpuRender` to the executable stanza of yahr.cabal.
-}
module GpuRender (renderGpu) where

import Control.Monad (when)
import Data.Word (Word32, Word64)
import Foreign
import Foreign.C.String (CString, peekCString)
import Foreign.C.Types
import qualified Data.Map as Map
import qualified Data.Vector.Storable as VS
import Codec.Picture (Image (..), PixelRGBF)

import Vectors
import qualified Scene as S
import qualified Cameras as C
import qualified Culling
import qualified Integrators as I
import qualified Lights as L

-- Opaque handle
data YahrScene

-- struct yahr_scene_desc / yahr_camera are marshalled by hand below (see include/yahr_b200.h
-- for the field order); sizes for x86-64 SysV.
foreign import ccall safe "yahr_b200_scene_create"
  c_scene_create :: Ptr () -> Ptr (Ptr YahrScene) -> IO CInt
foreign import ccall safe "yahr_b200_scene_destroy"
  c_scene_destroy :: Ptr YahrScene -> IO ()
-- `safe`: the call runs for a whole frame and the binary is -threaded (yahr.cabal:30)
foreign import ccall safe "yahr_b200_render"
  c_render :: Ptr YahrScene -> Ptr () -> CInt -> CInt -> Word64
           -> Ptr Float -> Ptr Word32 -> Ptr () -> IO CInt
-- The same frame with JuicyPixels' ImageRGBF -> RGB8 conversion of savePngImage (main.hs:142) applied on the GPU:
-- the buffer is an `Image PixelRGB8` payload (row-major, RGB interleaved), a quarter of the bytes cross PCIe.
foreign import ccall safe "yahr_b200_render_rgb8"
  c_render_rgb8 :: Ptr YahrScene -> Ptr () -> CInt -> CInt -> Word64 -> Ptr Word8 -> Ptr () -> IO CInt
foreign import ccall unsafe "yahr_b200_last_error"
  c_last_error :: IO CString

check :: String -> CInt -> IO ()
check what rc = when (rc /= 0) $ do
  msg <- c_last_error >>= peekCString
  ioError (userError (what ++ " failed (" ++ show rc ++ "): " ++ msg))

vec3s :: [Vec3] -> VS.Vector Float
vec3s vs = VS.fromList (concat [[x, y, z] | Vec3 x y z <- vs])

-- | Drop-in for `render`/`renderEval`/`renderPar` + `samplesToImage` (main.hs:68-107).
renderGpu :: S.Scene -> IO (Image PixelRGBF)
renderGpu scene = do
  let objs = S.objects scene >>= S.expand                      -- main.hs:44, primitive-ID order
      mats = S.materials scene
      matIndex = Map.fromList (zip (map S.id mats) [0 :: Word32 ..])
      midx o = matIndex Map.! S.materialId o                   -- same failure mode as main.hs:55
      tris = [o | o@S.Triangle {} <- objs]
      sphs = [o | o@S.Sphere {} <- objs]
      -- order entry: (kind << 31) | index-within-kind
      order = go 0 0 objs
        where go _ _ [] = []
              go s t (S.Sphere {} : r) = s : go (s + 1) t r
              go s t (S.Triangle {} : r) = (0x80000000 + t) : go s (t + 1) r
              go s t (_ : r) = go s t r
      Culling.BVH maxDepth split = S.cullingMode scene
      splitMode = case split of { Culling.Midpoint -> 0; Culling.SurfaceAreaHeuristic -> 1 } :: CInt
      cam = S.camera scene
      w = floor (C.imW cam) :: Int                             -- main.hs:122-123
      h = floor (C.imH cam) :: Int
      f32 = VS.unsafeWith :: VS.Vector Float -> (Ptr Float -> IO a) -> IO a
      u32 = VS.unsafeWith :: VS.Vector Word32 -> (Ptr Word32 -> IO a) -> IO a
  f32 (vec3s (map S.p0 tris)) $ \p0 -> f32 (vec3s (map S.p1 tris)) $ \p1 ->
   f32 (vec3s (map S.p2 tris)) $ \p2 -> f32 (vec3s (map S.n0 tris)) $ \n0 ->
   f32 (vec3s (map S.n1 tris)) $ \n1 -> f32 (vec3s (map S.n2 tris)) $ \n2 ->
   u32 (VS.fromList (map midx tris)) $ \tmat ->
   f32 (vec3s (map S.position sphs)) $ \sc -> f32 (VS.fromList (map S.radius sphs)) $ \sr ->
   u32 (VS.fromList (map midx sphs)) $ \smat -> u32 (VS.fromList order) $ \ord ->
   f32 (VS.fromList (concat [ [dr, dg, db, sr', sg, sb, sh]
                            | S.BlinnPhongMaterial _ _ (Vec3 dr dg db) (Vec3 sr' sg sb) sh <- mats ])) $ \mp ->
   f32 (VS.fromList (concat [ [px, py, pz, r, g, b]
                            | L.PointLight (Vec3 px py pz) (Vec3 r g b) <- S.lights scene ])) $ \lp ->
   allocaBytes 160 $ \desc -> allocaBytes 48 $ \cptr -> alloca $ \hptr -> do
     -- struct yahr_scene_desc (offsets: see include/yahr_b200.h)
     pokeByteOff desc 0 (fromIntegral (length tris) :: Word32)
     mapM_ (\(o, p) -> pokeByteOff desc o p) (zip [8, 16 ..] [p0, p1, p2, n0, n1, n2])
     pokeByteOff desc 56 tmat
     pokeByteOff desc 64 (fromIntegral (length sphs) :: Word32)
     pokeByteOff desc 72 sc; pokeByteOff desc 80 sr; pokeByteOff desc 88 smat
     pokeByteOff desc 96 ord
     pokeByteOff desc 104 (fromIntegral (length mats) :: Word32); pokeByteOff desc 112 mp
     pokeByteOff desc 120 (fromIntegral (length (S.lights scene)) :: Word32); pokeByteOff desc 128 lp
     pokeByteOff desc 136 (fromIntegral maxDepth :: CInt)
     pokeByteOff desc 140 splitMode
     pokeByteOff desc 144 (0 :: Word32); pokeByteOff desc 152 nullPtr   -- n_area_lights / area_lights: extension, unused by yahr
     do
       -- struct yahr_camera: imW imH focalLength lookDir[3] upDir[3] position[3]
       let Vec3 lx ly lz = C.lookDir cam; Vec3 ux uy uz = C.upDir cam; Vec3 qx qy qz = C.position cam
       pokeArray (castPtr cptr) [C.imW cam, C.imH cam, C.focalLength cam, lx, ly, lz, ux, uy, uz, qx, qy, qz]
       c_scene_create desc hptr >>= check "yahr_b200_scene_create"
       hdl <- peek hptr
       frame <- mallocForeignPtrArray (w * h * 3) :: IO (ForeignPtr Float)
       rc <- withForeignPtr frame $ \out ->
               c_render hdl cptr (fromIntegral (I.recursionDepth (S.integrator scene))) 1 0 out nullPtr nullPtr
       c_scene_destroy hdl
       check "yahr_b200_render" rc
       return (Image w h (VS.unsafeFromForeignPtr0 frame (w * h * 3)))
