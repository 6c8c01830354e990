"""Stitcher -- the reference's sequencing / alignment / mosaic class (Stitcher.py:14-525) on the B200 path.

Call surface kept: class attributes (isColorMode, direction, directIncre, fuseMethod, phaseResponseThreshold,
tempImageFeature, imageFusion + everything inherited from Method), the bound offset methods passed as callbacks
(Main.py:20), flowStitch / flowStitchWithMutiple / imageSetStitch / imageSetStitchWithMutiple, getStitchByOffset, fuseImage.
What changed underneath:
  * calculateOffsetForFeatureSearchIncre evaluates each (ROI size, direction) candidate with ONE fused device call
    (2 x SURF -> match -> vote, gpu.align_batch); flowStitch evaluates the predicted candidate of ALL pairs of a sequence
    in one batch and then replays the reference's search order on the host, falling back to single evaluations where the
    prediction failed -- candidate results are pure functions of (pair, i, direction), so the replay returns exactly the
    offsets the sequential loop would;
  * getStitchByOffset keeps the reference's integer bookkeeping on the host and runs paste + blend on a device canvas;
  * paths: Windows separators, unsorted glob and case-sensitive extensions are normalised.
"""
import glob
import os
import time

import cv2
import numpy as np

from . import gpu
from . import ImageFusion
from . import ImageUtility as Utility
from . import phase_wrap
from . import sharding


class ImageFeature():
    """Cache of the second image's features for the full-frame method (Stitcher.py:14-18)."""
    isBreak = True
    kps = None
    feature = None


def _norm_path(p):
    return p.replace("\\", os.sep)


def _list_images(directory, extension):
    """Sorted, case-insensitive `*.ext` listing (the reference's glob is unsorted and case-sensitive)."""
    ext = "." + extension.lower()
    try:
        names = [n for n in os.listdir(directory) if n.lower().endswith(ext)]
    except FileNotFoundError:
        return []
    return [os.path.join(directory, n) for n in sorted(names)]


def _imread(path, flag):
    return cv2.imdecode(np.fromfile(path, dtype=np.uint8), flag)


def _imwrite(path, image, encoder="b200"):
    """cv2.imwrite(path, image) (Stitcher.py:130-131, :196-197).  .jpg / .jpeg outputs are encoded by the library on the device --
    the file is byte-identical to cv2's (quality 95, 4:2:0, Annex-K tables); every other format, or encoder "cv2", goes through cv2."""
    img = np.asarray(image)
    ext = os.path.splitext(path)[1].lower()
    if encoder == "b200" and ext in (".jpg", ".jpeg") and img.dtype == np.uint8 and img.size > 0 \
            and (img.ndim == 2 or (img.ndim == 3 and img.shape[2] == 3)) and max(img.shape[:2]) <= 65535:
        data = gpu.jpeg_encode(img)
        try:
            with open(path, "wb") as f:
                f.write(data)
        except OSError:
            return False                      # cv2.imwrite reports an unwritable path by returning False
        return True
    return cv2.imwrite(path, image)


def _imread_gray_many(paths, decoder="b200"):
    """Grayscale decode of a tile sequence (Stitcher.py:68-69 decodes them one by one, twice).  JPEG files go through the
    library in batches -- host cores do the entropy decoding of many files at once, the device the IDCT; the pixels are
    bit-identical to cv2.imdecode(..., 0) -- anything the library does not support (progressive JPEG, PNG, ...) through cv2."""
    datas = [np.fromfile(p, dtype=np.uint8) for p in paths]
    images = [None] * len(paths)
    if decoder == "b200":
        groups = {}
        for k, d in enumerate(datas):
            if d.size > 3 and d[0] == 0xFF and d[1] == 0xD8:
                try:
                    groups.setdefault(gpu.jpeg_info(d)[:2], []).append(k)
                except gpu.JpegUnsupported:
                    pass
        for idx in groups.values():
            try:
                out = gpu.jpeg_decode_gray([datas[k] for k in idx])
            except gpu.JpegUnsupported:
                continue
            for j, k in enumerate(idx):
                images[k] = out[j]
    for k, d in enumerate(datas):
        if images[k] is None:
            images[k] = cv2.imdecode(d, cv2.IMREAD_GRAYSCALE)
    return images


def _imread_color_many(paths, decoder="b200"):
    """cv2.imdecode(..., IMREAD_COLOR) of a tile sequence (Stitcher.py:382,401), JPEG files batched through the library."""
    datas = [np.fromfile(p, dtype=np.uint8) for p in paths]
    images = [None] * len(paths)
    if decoder == "b200":
        groups = {}
        for k, d in enumerate(datas):
            try:
                groups.setdefault(gpu.jpeg_info(d)[:2], []).append(k)
            except gpu.VfsmsError:
                pass
        for idx in groups.values():
            for c0 in range(0, len(idx), 16):                 # bounded host staging: 16 colour tiles per call
                part = idx[c0:c0 + 16]
                try:
                    out = gpu.jpeg_decode_bgr([datas[k] for k in part])
                except gpu.JpegUnsupported:
                    continue
                for j, k in enumerate(part):
                    images[k] = out[j]
    for k, d in enumerate(datas):
        if images[k] is None:
            images[k] = cv2.imdecode(d, cv2.IMREAD_COLOR)
    return images


_last_sequence_has_color = False     # set by _load_sequence: the colour twin of the stack holds the sequence


def _load_sequence(paths, decoder="b200", keep_on_device=True, color=False):
    """Decode a tile sequence ONCE.  -> (host gray images, on_device).  With decoder "b200" and equally sized tiles the
    tiles also stay in the library's device-resident stack (slot k = paths[k]): JPEG files are decoded straight into it,
    the rest is decoded by cv2 and uploaded; the host copies are then a download of the stack.  color: the colour twin of the
    stack is filled as well -- JPEG files by the SAME entropy-decoding pass that yields the gray tile (the reference decodes each
    file in gray at Stitcher.py:68-69 and again in colour at :382 / :401)."""
    if decoder != "b200" or not keep_on_device:
        return _imread_gray_many(paths, decoder), False
    datas = [np.fromfile(p, dtype=np.uint8) for p in paths]
    shapes, jpeg_ok = [], []
    for d in datas:
        try:
            r, c, _ = gpu.jpeg_info(d)
            shapes.append((r, c)); jpeg_ok.append(True)
        except gpu.VfsmsError:
            shapes.append(None); jpeg_ok.append(False)
    others = {k: cv2.imdecode(datas[k], cv2.IMREAD_GRAYSCALE) for k in range(len(paths)) if not jpeg_ok[k]}
    for k, im in others.items():
        shapes[k] = None if im is None else im.shape[:2]
    if any(sh is None for sh in shapes) or len(set(shapes)) != 1:
        return _imread_gray_many(paths, decoder), False
    rows, cols = shapes[0]
    n = len(paths)

    def fill(with_color):
        gpu.tiles_reserve(n, rows, cols)
        k = 0
        while k < n:                                 # runs of JPEG files are decoded with one call each
            e = k
            while e < n and jpeg_ok[e] == jpeg_ok[k]:
                e += 1
            if jpeg_ok[k]:
                (gpu.tiles_decode_jpeg_bgr if with_color else gpu.tiles_decode_jpeg)(k, [datas[j] for j in range(k, e)])
            else:
                gpu.tiles_upload(k, np.stack([others[j] for j in range(k, e)]))
                if with_color:
                    gpu.tiles_upload_bgr(k, np.stack([cv2.imdecode(datas[j], cv2.IMREAD_COLOR) for j in range(k, e)]))
            k = e
    global _last_sequence_has_color
    _last_sequence_has_color = bool(color)
    try:
        fill(color)
    except gpu.VfsmsError:
        # colour: the twin did not fit (3 x the gray stack) -> gray stack only, the mosaic then decodes the colour tiles batch by batch.
        # gray (or the retry): no room for the stack, or a file the parser accepted turned out undecodable -> host tiles, per-file decode
        _last_sequence_has_color = False
        try:
            if not color:
                raise
            fill(False)
        except gpu.VfsmsError:
            return _imread_gray_many(paths, decoder), False
    host = gpu.tiles_download(0, n, rows, cols)
    return [host[j] for j in range(n)], True


class Stitcher(Utility.Method):
    isColorMode = True
    direction = 1           # 1: B below A, 2: B right of A, 3: B above A, 4: B left of A
    directIncre = 1         # 1, 0 or -1
    fuseMethod = "notFuse"
    phaseResponseThreshold = 0.15
    phaseMode = "reference"  # "reference": Stitcher.py:205-258 verbatim in behaviour (wrong by construction);
                             # "wrapAware": corrected sign, aliases resolved by overlap ZNCC (phase_wrap.py)
    phaseAcceptZncc = 0.5    # wrapAware: a candidate must correlate at least this well over the pixels the ROIs would share
    tempImageFeature = ImageFeature()
    imageFusion = ImageFusion.ImageFusion()
    batchPairs = 16         # pairs evaluated per fused device call in flowStitch
    decoder = "b200"        # "b200": JPEG tiles decoded by the library (bit-identical to cv2); "cv2": cv2.imdecode
    encoder = "b200"        # "b200": .jpg results encoded by the library (file byte-identical to cv2.imwrite's); "cv2": cv2.imwrite

    # ------------------------------------------------------------------ direction bookkeeping
    def directionIncrease(self, direction):
        """Stitcher.py:36-47."""
        direction += self.directIncre
        if direction == 5:
            direction = 1
        if direction == 0:
            direction = 4
        return direction

    # ------------------------------------------------------------------ sequencing
    def flowStitch(self, fileList, caculateOffsetMethod):
        """Stitcher.py:49-94.  Returns ((status, endfileIndex), stitchImage)."""
        self.printAndWrite("Stitching the directory which have " + str(fileList[0]))
        fileNum = len(fileList)
        offsetList = []
        describtion = ""
        startTime = time.time()
        status = True
        endfileIndex = 0
        batched = self._is_incre_feature_method(caculateOffsetMethod) and self.featureMethod == "surf" \
            and self.offsetCaculate == "mode" and not self.isEnhance
        # decoded once (the reference decodes every tile twice, and a third time for the mosaic)
        images, self._on_device = _load_sequence(fileList, self.decoder, keep_on_device=True, color=bool(self.isColorMode))
        self._stack_color = bool(self.isColorMode) and self._on_device and _last_sequence_has_color
        table = {}
        for fileIndex in range(0, fileNum - 1):
            self.printAndWrite("stitching " + str(fileList[fileIndex]) + " and " + str(fileList[fileIndex + 1]))
            if batched:
                if (fileIndex, 1, self.direction) not in table:
                    self._prefetch_candidates(images, fileIndex, table)
                (status, offset) = self._incre_search([images[fileIndex], images[fileIndex + 1]], self._surf_evaluator(fileIndex, table))
            else:
                (status, offset) = caculateOffsetMethod([images[fileIndex], images[fileIndex + 1]])
            if status == False:
                describtion = "  " + str(fileList[fileIndex]) + " and " + str(fileList[fileIndex + 1]) + " can not be stitched"
                break
            offsetList.append(offset)
            endfileIndex = fileIndex + 1
        endTime = time.time()
        self.printAndWrite("The time of registering is " + str(endTime - startTime) + "s")

        self.printAndWrite("start stitching")
        startTime = time.time()
        self._gray_cache = dict(zip(fileList, images))      # gray mosaics reuse the decoded tiles (the reference decodes them again)
        self._stack_files = list(fileList) if self._on_device else None
        try:
            stitchImage = self.getStitchByOffset(fileList, offsetList)
        finally:
            self._gray_cache = {}
            self._stack_files = None
            self._on_device = False
            self._stack_color = False
        endTime = time.time()
        self.printAndWrite("The time of fusing is " + str(endTime - startTime) + "s")
        if status == False:
            self.printAndWrite(describtion)
        return ((status, endfileIndex), stitchImage)

    def flowStitchWithMutiple(self, fileList, caculateOffsetMethod):
        """Segments a sequence at every failed pair (Stitcher.py:96-127)."""
        result = []
        totalNum = len(fileList)
        startNum = 0
        while 1:
            (status, stitchResult) = self.flowStitch(fileList[startNum:totalNum], caculateOffsetMethod)
            result.append(stitchResult)
            self.tempImageFeature.isBreak = True
            startNum = startNum + status[1] + 1
            if startNum == totalNum:
                break
            if startNum == (totalNum - 1):
                flag = cv2.IMREAD_COLOR if self.isColorMode else cv2.IMREAD_GRAYSCALE
                result.append(_imread(fileList[startNum], flag))
                break
            self.printAndWrite("stitching Break, start from " + str(fileList[startNum]) + " again")
        return result

    def imageSetStitch(self, projectAddress, outputAddress, fileNum, caculateOffsetMethod, startNum=1, fileExtension="jpg",
                       outputfileExtension="jpg"):
        """Stitcher.py:129-151."""
        projectAddress = _norm_path(projectAddress); outputAddress = _norm_path(outputAddress)
        for i in range(startNum, fileNum + 1):
            fileList = _list_images(os.path.join(projectAddress, str(i)), fileExtension)
            os.makedirs(outputAddress, exist_ok=True)
            Stitcher.outputAddress = outputAddress
            (status, result) = self.flowStitch(fileList, caculateOffsetMethod)
            self.tempImageFeature.isBreak = True
            _imwrite(os.path.join(outputAddress, "stitching_result_" + str(i) + "." + outputfileExtension), result, self.encoder)
            if status[0] == False:
                self.printAndWrite("stitching Failed")

    def imageSetStitchWithMutiple(self, projectAddress, outputAddress, fileNum, caculateOffsetMethod, startNum=1,
                                  fileExtension="jpg", outputfileExtension="jpg"):
        """Stitcher.py:153-182 (the entry Main.py calls)."""
        projectAddress = _norm_path(projectAddress); outputAddress = _norm_path(outputAddress)
        for i in range(startNum, fileNum + 1):
            startTime = time.time()
            fileAddress = os.path.join(projectAddress, str(i))
            fileList = _list_images(fileAddress, fileExtension)
            os.makedirs(outputAddress, exist_ok=True)
            Stitcher.outputAddress = outputAddress
            result = self.flowStitchWithMutiple(fileList, caculateOffsetMethod)
            self.tempImageFeature.isBreak = True
            if len(result) == 1:
                _imwrite(os.path.join(outputAddress, "stitching_result_" + str(i) + "." + outputfileExtension), result[0], self.encoder)
            else:
                for j in range(0, len(result)):
                    _imwrite(os.path.join(outputAddress, "stitching_result_" + str(i) + "_" + str(j + 1) + "." + outputfileExtension), result[j], self.encoder)
            endTime = time.time()
            print("Time Consuming for " + fileAddress + " is " + str(endTime - startTime))

    # ------------------------------------------------------------------ alignment
    def calculateOffsetForPhaseCorrleate(self, dirAddress):
        """Dead in the reference: it dereferences a `self.phase` that is never created (Stitcher.py:184-203)."""
        raise AttributeError("'Stitcher' object has no attribute 'phase' (calculateOffsetForPhaseCorrleate is dead code in the reference)")

    def _roi_origin_back(self, offset, images, i, localDirection):
        """Add the ROI origin back (Stitcher.py:243-251, 353-360)."""
        (imageA, imageB) = images
        if localDirection == 1:
            offset[0] = offset[0] + imageA.shape[0] - int(i * self.roiRatio * imageA.shape[0])
        elif localDirection == 2:
            offset[1] = offset[1] + imageA.shape[1] - int(i * self.roiRatio * imageA.shape[1])
        elif localDirection == 3:
            offset[0] = offset[0] - (imageB.shape[0] - int(i * self.roiRatio * imageB.shape[0]))
        elif localDirection == 4:
            offset[1] = offset[1] - (imageB.shape[1] - int(i * self.roiRatio * imageB.shape[1]))
        return offset

    def _incre_search(self, images, evaluate):
        """The ROI-size x direction search loop shared by the incremental methods (Stitcher.py:316-367, 215-258).
        evaluate(i, direction) -> (status, [dRow, dCol]) for one candidate."""
        self._cur_images = images          # the table-backed evaluator of flowStitch fetches ROIs on demand
        offset = [0, 0]
        status = False
        maxI = int(np.floor(0.5 / self.roiRatio) + 1) + 1
        iniDirection = self.direction
        localDirection = iniDirection
        for i in range(1, maxI):
            while True:
                (status, offset) = evaluate(i, localDirection)
                if status:
                    break
                localDirection = self.directionIncrease(localDirection)
                if localDirection == iniDirection:
                    break
            if status:
                offset = self._roi_origin_back(list(offset), images, i, localDirection)
                self.direction = localDirection
                break
        if status == False:
            return (status, "  The two images can not match")
        self.printAndWrite("  The offset of stitching: dx is " + str(offset[0]) + " dy is " + str(offset[1]))
        return (status, offset)

    def calculateOffsetForPhaseCorrleateIncre(self, images):
        """Incremental-ROI phase correlation (Stitcher.py:205-258); accepts when response > phaseResponseThreshold."""
        (imageA, imageB) = images

        def evaluate(i, d):
            roiA = self.getROIRegionForIncreMethod(imageA, direction=d, order="first", searchRatio=i * self.roiRatio)
            roiB = self.getROIRegionForIncreMethod(imageB, direction=d, order="second", searchRatio=i * self.roiRatio)
            if self.phaseMode == "wrapAware":
                (status, offset, _, _) = phase_wrap.resolve(roiA, roiB, gpu.phase_correlate, gpu.overlap_sums, accept=self.phaseAcceptZncc)
                return (status, offset)
            (shift, response) = gpu.phase_correlate(roiA, roiB)
            return (response > self.phaseResponseThreshold, [int(shift[1]), int(shift[0])])       # truncation, Stitcher.py:231-232
        return self._incre_search(images, evaluate)

    def _enhance(self, image):
        """Optional pre-processing (Stitcher.py:269-276, 327-334): CLAHE(clipLimit, tileSize) or equalizeHist, on the device."""
        return gpu.enhance(image, clahe=bool(self.isClahe), clip_limit=self.clipLimit, tile_size=self.tileSize)

    def calculateOffsetForFeatureSearch(self, images):
        """Full-frame features with the B-feature cache (Stitcher.py:260-304)."""
        (imageA, imageB) = images
        offset = [0, 0]
        status = False
        if self.isEnhance:
            imageA = self._enhance(imageA); imageB = self._enhance(imageB)
        if self.tempImageFeature.isBreak:
            (kpsA, featuresA) = self.detectAndDescribe(imageA, featureMethod=self.featureMethod)
        else:
            kpsA = self.tempImageFeature.kps
            featuresA = self.tempImageFeature.feature
        (kpsB, featuresB) = self.detectAndDescribe(imageB, featureMethod=self.featureMethod)
        self.tempImageFeature.isBreak = False
        self.tempImageFeature.kps = kpsB
        self.tempImageFeature.feature = featuresB
        if featuresA is not None and featuresB is not None and len(featuresA) and len(featuresB):
            matches = self.matchDescriptors(featuresA, featuresB)
            if self.offsetCaculate == "mode":
                (status, offset) = self.getOffsetByMode(kpsA, kpsB, matches, offsetEvaluate=self.offsetEvaluate)
            elif self.offsetCaculate == "ransac":
                (status, offset, adjustH) = self.getOffsetByRansac(kpsA, kpsB, matches, offsetEvaluate=self.offsetEvaluate)
        if status == False:
            self.tempImageFeature.isBreak = True
            return (status, "  The two images can not match")
        self.tempImageFeature.isBreak = False
        self.printAndWrite("  The offset of stitching: dx is " + str(offset[0]) + " dy is " + str(offset[1]))
        return (status, offset)

    def _candidate_staged(self, roiA, roiB):
        """One search-loop body through the three-stage API (any featureMethod / offsetCaculate)."""
        status, offset = False, [0, 0]
        kpsA, featuresA = self.detectAndDescribe(roiA, featureMethod=self.featureMethod)
        kpsB, featuresB = self.detectAndDescribe(roiB, featureMethod=self.featureMethod)
        if featuresA is not None and featuresB is not None and len(featuresA) and len(featuresB):
            matches = self.matchDescriptors(featuresA, featuresB)
            if self.offsetCaculate == "mode":
                (status, offset) = self.getOffsetByMode(kpsA, kpsB, matches, offsetEvaluate=self.offsetEvaluate)
            elif self.offsetCaculate == "ransac":
                (status, offset, adjustH) = self.getOffsetByRansac(kpsA, kpsB, matches, offsetEvaluate=self.offsetEvaluate)
        return (status, offset)

    def calculateOffsetForFeatureSearchIncre(self, images):
        """Incremental-ROI feature search (Stitcher.py:306-367)."""
        (imageA, imageB) = images
        fused = self.featureMethod == "surf" and self.offsetCaculate == "mode" and not self.isEnhance
        params = self._surf_params() if fused else None

        def evaluate(i, d):
            roiA = self.getROIRegionForIncreMethod(imageA, direction=d, order="first", searchRatio=i * self.roiRatio)
            roiB = self.getROIRegionForIncreMethod(imageB, direction=d, order="second", searchRatio=i * self.roiRatio)
            if fused:
                r = gpu.align_batch(roiA[None], roiB[None], params=params, ratio=self.searchRatio, offset_evaluate=self.offsetEvaluate)[0]
                return (bool(r["status"]), [int(r["d_row"]), int(r["d_col"])])
            if self.isEnhance:
                roiA = self._enhance(roiA); roiB = self._enhance(roiB)
            return self._candidate_staged(roiA, roiB)
        return self._incre_search(images, evaluate)

    # -- batched evaluation used by flowStitch ---------------------------------------------------------------------
    def _is_incre_feature_method(self, method):
        return getattr(method, "__func__", None) is Stitcher.calculateOffsetForFeatureSearchIncre and getattr(method, "__self__", None) is self

    def _prefetch_candidates(self, images, start, table):
        """Evaluate candidate (i = 1, predicted direction) of pairs start .. start+batchPairs-1 with one device call."""
        d = self.direction
        idx = []
        for k in range(start, min(len(images) - 1, start + self.batchPairs)):
            if images[k].shape == images[start].shape and images[k + 1].shape == images[start].shape:
                idx.append(k)
            else:
                break
        if not idx:
            return
        if getattr(self, "_on_device", False):
            L = int(np.floor((images[start].shape[0] if d in (1, 3) else images[start].shape[1]) * self.roiRatio))
            res = gpu.tiles_align(idx[0], len(idx), d, L, params=self._surf_params(), ratio=self.searchRatio, offset_evaluate=self.offsetEvaluate)
            for k, r in zip(idx, res):
                table[(k, 1, d)] = (bool(r["status"]), [int(r["d_row"]), int(r["d_col"])])
            return
        roisA = np.stack([np.ascontiguousarray(self.getROIRegionForIncreMethod(images[k], d, "first", self.roiRatio)) for k in idx])
        roisB = np.stack([np.ascontiguousarray(self.getROIRegionForIncreMethod(images[k + 1], d, "second", self.roiRatio)) for k in idx])
        res = gpu.align_batch(roisA, roisB, params=self._surf_params(), ratio=self.searchRatio, offset_evaluate=self.offsetEvaluate)
        for k, r in zip(idx, res):
            table[(k, 1, d)] = (bool(r["status"]), [int(r["d_row"]), int(r["d_col"])])

    def _surf_evaluator(self, k, table):
        def evaluate(i, d):
            key = (k, i, d)
            if key not in table and getattr(self, "_on_device", False):
                shape = self._cur_images[0].shape
                L = int(np.floor((shape[0] if d in (1, 3) else shape[1]) * (i * self.roiRatio)))
                r = gpu.tiles_align(k, 1, d, L, params=self._surf_params(), ratio=self.searchRatio, offset_evaluate=self.offsetEvaluate)[0]
                table[key] = (bool(r["status"]), [int(r["d_row"]), int(r["d_col"])])
            if key not in table:
                roiA = self.getROIRegionForIncreMethod(self._cur_images[0], direction=d, order="first", searchRatio=i * self.roiRatio)
                roiB = self.getROIRegionForIncreMethod(self._cur_images[1], direction=d, order="second", searchRatio=i * self.roiRatio)
                r = gpu.align_batch(roiA[None], roiB[None], params=self._surf_params(), ratio=self.searchRatio,
                                    offset_evaluate=self.offsetEvaluate)[0]
                table[key] = (bool(r["status"]), [int(r["d_row"]), int(r["d_col"])])
            st, off = table[key]
            return (st, list(off))
        return evaluate

    # ------------------------------------------------------------------ mosaic
    def getStitchByOffset(self, fileList, originOffsetList):
        """Global offset rectification on the host (Stitcher.py:378-431, integer bookkeeping kept verbatim in behaviour),
        paste + blend on a device canvas (Stitcher.py:433-486)."""
        flag = cv2.IMREAD_COLOR if self.isColorMode else cv2.IMREAD_GRAYSCALE
        originOffsetList.insert(0, [0, 0])          # the reference mutates its argument the same way
        n = len(originOffsetList)
        stack_files = getattr(self, "_stack_files", None)
        if self.isColorMode and self.fuseMethod in gpu.DEVICE_MOSAIC_METHODS and getattr(self, "_stack_color", False) \
                and stack_files is not None and stack_files[:n] == list(fileList[:n]):
            # colour tiles are already in HBM (decoded together with the gray ones): no second decode, no upload
            shape = next(iter(self._gray_cache.values())).shape[:2]
            origins, rois, (resultRow, resultCol) = sharding.rectify_offsets(originOffsetList, [shape] * n)
            self.printAndWrite("  The rectified offsetList is " + str([[int(o[0]), int(o[1])] for o in origins]))
            return gpu.tiles_mosaic_bgr(0, n, np.asarray(origins, np.int32), rois, np.asarray(originOffsetList, np.int32), self.fuseMethod,
                                        (resultRow, resultCol))
        cache = {} if self.isColorMode else getattr(self, "_gray_cache", {})
        if self.isColorMode and self.decoder == "b200":        # colour tiles: one batched decode instead of a cv2 call per tile
            cache = dict(zip(fileList[:n], _imread_color_many(fileList[:n], self.decoder)))

        def _imread(path, flag):
            return cache[path] if path in cache else globals()["_imread"](path, flag)
        imageList = [_imread(fileList[i], flag) for i in range(n)]
        origins, rois, (resultRow, resultCol) = sharding.rectify_offsets(originOffsetList, [im.shape[:2] for im in imageList])
        offsetList = [[int(o[0]), int(o[1])] for o in origins]
        self.printAndWrite("  The rectified offsetList is " + str(offsetList))
        if self.fuseMethod in ("multiBandBlending", "optimalSeamLine"):
            assert self.isColorMode is False, "The %s is not support for color mode in this code" % self.fuseMethod

        same_shape = all(im.shape == imageList[0].shape for im in imageList)
        if same_shape and self.fuseMethod in gpu.DEVICE_MOSAIC_METHODS and not self.isColorMode \
                and getattr(self, "_stack_files", None) is not None and self._stack_files[:n] == list(fileList[:n]):
            return gpu.tiles_mosaic(0, n, np.asarray(offsetList, np.int32), rois, np.asarray(originOffsetList, np.int32),
                                    self.fuseMethod, (resultRow, resultCol))
        if same_shape and self.fuseMethod in gpu.DEVICE_MOSAIC_METHODS:
            return gpu.mosaic(np.stack(imageList), np.asarray(offsetList, np.int32), rois, np.asarray(originOffsetList, np.int32),
                              self.fuseMethod, (resultRow, resultCol))
        # mixed tile sizes / multi-band: per-ROI device fuse on a host canvas with the reference's data model
        shape = (resultRow, resultCol, 3) if self.isColorMode else (resultRow, resultCol)
        stitchResult = np.zeros(shape, np.int16) - 1
        for i in range(0, n):
            self.printAndWrite("  stitching " + str(fileList[i]))
            r0, c0 = offsetList[i]
            h, w = imageList[i].shape[:2]
            if i == 0 or self.fuseMethod == "notFuse":
                stitchResult[r0:r0 + h, c0:c0 + w] = imageList[i]
                continue
            a0, b0, a1, b1 = (int(v) for v in rois[i])
            roiA = stitchResult[a0:a1, b0:b1].copy()
            stitchResult[r0:r0 + h, c0:c0 + w] = imageList[i]
            roiB = stitchResult[a0:a1, b0:b1].copy()
            stitchResult[a0:a1, b0:b1] = self.fuseImage([roiA, roiB], originOffsetList[i][0], originOffsetList[i][1])
        stitchResult[stitchResult == -1] = 0
        return stitchResult.astype(np.uint8)

    def fuseImage(self, images, dx, dy):
        """Blend dispatch (Stitcher.py:488-525); dx / dy = the ORIGINAL pair offset (rows / columns)."""
        self.imageFusion.isColorMode = self.isColorMode
        (imageA, imageB) = images
        if self.fuseMethod == "optimalSeamLine":
            assert self.isColorMode is False, "The optimal seam line is not support for color mode in this code"
            return self.imageFusion.fuseByOptimalSeamLine(images, self.direction)
        if self.fuseMethod == "multiBandBlending":
            assert self.isColorMode is False, "The multi Band Blending is not support for color mode in this code"
        if self.fuseMethod not in gpu.FUSE_METHODS:
            return np.zeros(np.asarray(imageA).shape, np.uint8)            # unknown method: the reference returns zeros
        return gpu.fuse_roi(imageA, imageB, self.fuseMethod, dx, dy)


if __name__ == "__main__":
    import sys
    st = Stitcher()
    a = cv2.imread(sys.argv[1], 0); b = cv2.imread(sys.argv[2], 0)
    print(st.calculateOffsetForFeatureSearchIncre([a, b]))
