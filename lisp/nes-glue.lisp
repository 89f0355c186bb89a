;;;; nes-glue.lisp -- load THIS instead of sparse-cholesky.lisp.
;;;;
;;;; Drop-in replacement of the reference's CHOLMOD binding (sparse-cholesky.lisp, 614 lines) over
;;;; libnes.so (include/nes.h): it defines every name the other files of pkhuong/cholesky-is-magic use
;;;; from that file -- *cholmod-common*, with-cholmod, the cholmod-get-/set- accessors, the alien types
;;;; cholmod-sparse / cholmod-factor, cholmod-copy-sparse, cholmod-free-sparse, cholmod-analyze,
;;;; cholmod-free-work, make-sparse-from-triplet-vector, scale-sparse!, scale-sparse,
;;;; solve-sparse-state (+ accessors), free-sparse-state, solve-sparse (-one-shot, -recycle),
;;;; solve-dense, sparse-m*, flush -- so that newton-solve.lisp, sparse-newton-solve.lisp,
;;;; affine-scaling.lisp and primal-dual-affine-scaling.lisp load UNCHANGED on top of it and their own
;;;; tests -- (test 30) of newton-solve.lisp:202-211 and sparse-newton-solve.lisp:260-269, test-m-n with
;;;; its leak check (sparse-newton-solve.lisp:239-258), pdas, affine-scaling -- run on the GPU.
;;;;
;;;; One function of the other files has to be redefined because it memcpy's the HOST value array of a
;;;; cholmod_sparse (device resident here): affine-A-copy (affine-scaling.lisp:29-35).  Its replacement
;;;; is at the end of this file under #+nes-after-affine-scaling; load that form after
;;;; affine-scaling.lisp (see "Load order" below).
;;;;
;;;; This file mirrors cholesky-is-magic_b200/nes.py and sparse_cholesky.py one to one (those are the
;;;; same stubs in Python ctypes, exercised by the test-suite).  It could NOT be executed where it was
;;;; written: the build image has no Common Lisp.  It uses only sb-alien / sb-sys operators the reference
;;;; itself uses, plus with-pinned-objects / vector-sap.
;;;;
;;;; Load order (SBCL, matlisp and alexandria available as for the reference):
;;;;   (defparameter cl-user::*libnes* "/path/to/cholesky-is-magic_b200/libnes.so")
;;;;   (load "standard-form.lisp") (load "read-mps.lisp")
;;;;   (load "lisp/nes-glue.lisp")                       ; instead of sparse-cholesky.lisp
;;;;   (load "newton-solve.lisp") (load "sparse-newton-solve.lisp")
;;;;   (load "affine-scaling.lisp") (load "primal-dual-affine-scaling.lisp")
;;;;   (pushnew :nes-after-affine-scaling *features*) (load "lisp/nes-glue.lisp")   ; the one override
;;;;   (test 30)                                          ; the reference's own property tests
;;;;
;;;; Semantics kept from the reference (file:line = sparse-cholesky.lisp unless another file is named):
;;;;   * status protocol: set-status 0 -> factorize -> get-status; non-zero => NIL (:418-421, :511-514,
;;;;     :541-544).  1 = not positive definite (column in nes-get-minor).
;;;;   * frees take the ADDRESS of the handle and return non-zero on success (:75-77, :145-147).
;;;;   * scale-sparse! scales columns (mode 2, :469); the matrix handle keeps s and every later product /
;;;;     factorization uses A diag(s) -- no value array is rewritten.
;;;;   * sparse-m* :y / :output aliasing (:567-614): y <- alpha op(A) x + beta y, result written into
;;;;     OUTPUT when given (may be Y itself, as clear-delta-z does, sparse-newton-solve.lisp:116-119).
;;;;   * solve-sparse-recycle returns its solution IN the right-hand side's matrix (:560).
;;;;   * the malloc-count / memory-inuse counters return to zero after the frees + cholmod-free-work.

#-nes-after-affine-scaling
(progn

(defvar cl-user::*libnes* "libnes.so")
(load-shared-object cl-user::*libnes*)

;;; ---- context: wrapper.c:8-16 + :40-43 ----------------------------------------------------------
(define-alien-type cholmod-common (struct nes-ctx))

(define-alien-routine ("nes_allocate" cholmod-allocate) (* cholmod-common))
(define-alien-routine ("nes_release" cholmod-release) void (ptr (* cholmod-common)))
(define-alien-routine ("nes_start" cholmod-start) int (ptr (* cholmod-common)))
(define-alien-routine ("nes_finish" cholmod-finish) int (ptr (* cholmod-common)))
(define-alien-routine ("nes_defaults" cholmod-defaults) int (ptr (* cholmod-common)))
(define-alien-routine ("nes_free_work" cholmod-free-work) int (ptr (* cholmod-common)))
(define-alien-routine ("nes_set_device" nes-set-device) int (ptr (* cholmod-common)) (device int))
(define-alien-routine ("nes_last_error" nes-last-error) c-string (ptr (* cholmod-common)))
(define-alien-routine ("nes_get_minor" nes-get-minor) int (ptr (* cholmod-common)))

;;; the 19 accessor pairs of :8-38 / wrapper.c:31-52, same Lisp names, C names nes_get_/nes_set_<field>
(macrolet ((def (suffix c-suffix type)
             `(progn
                (define-alien-routine (,(format nil "nes_get_~A" c-suffix)
                                       ,(alexandria:format-symbol *package* "CHOLMOD-GET-~A" suffix))
                  ,type (ptr (* cholmod-common)))
                (define-alien-routine (,(format nil "nes_set_~A" c-suffix)
                                       ,(alexandria:format-symbol *package* "CHOLMOD-SET-~A" suffix))
                  ,type (ptr (* cholmod-common)) (value ,type)))))
  (def #:print "print" int)
  (def #:print-function "print_function" (* t))
  (def #:dbound "dbound" double)
  (def #:supernodal-switch "supernodal_switch" double)
  (def #:supernodal "supernodal" int)
  (def #:selected "selected" int)
  (def #:itype "itype" int)
  (def #:dtype "dtype" int)
  (def #:status "status" int)
  (def #:fl "fl" double)
  (def #:lnz "lnz" double)
  (def #:anz "anz" double)
  (def #:modfl "modfl" double)
  (def #:malloc-count "malloc_count" size-t)
  (def #:memory-usage "memory_usage" size-t)
  (def #:memory-inuse "memory_inuse" size-t)
  (def #:rowfacfl "rowfacfl" double)
  (def #:aatfl "aatfl" double)
  (def #:blas-ok "blas_ok" int))

(defvar *cholmod-common*)

(define-alien-routine fflush void (handle (* t)))
(defun flush () (fflush nil))

(defun make-default-common ()                                   ; :389-394
  (let ((common (cholmod-allocate)))
    ;; nes_start binds the context to an sm_100 device and returns 0 when there is none:
    ;; there is no CPU fallback behind this library
    (unless (plusp (cholmod-start common))
      (let ((message (nes-last-error common)))
        (cholmod-release common)
        (error "nes_start failed: ~A" message)))
    (cholmod-defaults common)
    common))

(defun free-common (common)                                     ; :396-399
  (cholmod-finish common)
  (cholmod-release common)
  nil)

(defmacro with-cholmod (() &body body)                          ; :400-406, unchanged
  (let ((temp (gensym "COMMON")))
    `(let* ((,temp (make-default-common))
            (*cholmod-common* ,temp))
       (unwind-protect
            (locally ,@body)
         (free-common ,temp)))))

;;; ---- matrices ----------------------------------------------------------------------------------
;;; A (* cholmod-sparse) is a nes_matrix handle: opaque except for its public prefix, the leading
;;; fields of cholmod_sparse (:45-48), so (slot A 'nrow) etc. of the other files keep working.
(define-alien-type cholmod-sparse
  (struct nes-matrix-header
          (nrow size-t)
          (ncol size-t)
          (nzmax size-t)))
(define-alien-type cholmod-factor (struct nes-factor))

(define-alien-routine ("nes_dense_to_matrix" nes-dense-to-matrix) (* cholmod-sparse)
  (a (* double)) (nrow size-t) (ncol size-t) (ld size-t) (common (* cholmod-common)))
(define-alien-routine ("nes_triplet_to_sparse" nes-triplet-to-sparse) (* cholmod-sparse)
  (row (* int)) (col (* int)) (val (* double)) (nnz size-t) (nrow size-t) (ncol size-t)
  (common (* cholmod-common)))
(define-alien-routine ("nes_copy_matrix" cholmod-copy-sparse) (* cholmod-sparse)      ; :128-130
  (a (* cholmod-sparse)) (common (* cholmod-common)))
(define-alien-routine ("nes_free_matrix" cholmod-free-sparse) int                      ; :75-77
  (a (* (* cholmod-sparse))) (common (* cholmod-common)))
(define-alien-routine ("nes_matrix_nnz" nes-matrix-nnz) size-t (a (* cholmod-sparse)))
(define-alien-routine ("nes_scale" nes-scale) int                                      ; cholmod_scale :329-333
  (s (* double)) (scale int) (a (* cholmod-sparse)) (common (* cholmod-common)))
(define-alien-routine ("nes_unscale" nes-unscale) int
  (a (* cholmod-sparse)) (common (* cholmod-common)))
(define-alien-routine ("nes_sdmult" nes-sdmult) int                                    ; cholmod_sdmult :335-342
  (a (* cholmod-sparse)) (transpose int) (alpha (* double)) (beta (* double))
  (x (* double)) (y (* double)) (common (* cholmod-common)))

(defun cholmod-nnz (sparse common)                                                     ; :84-86
  (declare (ignore common))
  (nes-matrix-nnz sparse))

(defmacro with-store-saps ((&rest bindings) &body body)
  "Bind each VAR to the address of the double-float store of a matlisp matrix, pinned for BODY.
The library reads / writes host memory only during the call, as CHOLMOD's copies did (:357-361)."
  (let ((stores (loop for nil in bindings collect (gensym "STORE"))))
    `(let ,(loop for (nil matrix) in bindings
                 for store in stores
                 collect `(,store (the (simple-array double-float 1) (matlisp::store ,matrix))))
       (sb-sys:with-pinned-objects ,stores
         (let ,(loop for (var nil) in bindings
                     for store in stores
                     collect `(,var (sb-alien:sap-alien (sb-sys:vector-sap ,store) (* double))))
           ,@body)))))

(defun make-sparse-from-triplet-vector (nrow ncol vector        ; :433-459
                                        &aux (nnz (length vector))
                                          (common *cholmod-common*))
  ;; duplicates are summed and columns sorted by the library (cholmod_triplet_to_sparse + cholmod_sort)
  (let ((rows (make-array nnz :element-type '(signed-byte 32)))
        (cols (make-array nnz :element-type '(signed-byte 32)))
        (xs   (make-array nnz :element-type 'double-float)))
    (loop for i upfrom 0
          for triplet across vector
          do (assert (< (triplet-row triplet) nrow))
             (assert (< (triplet-col triplet) ncol))
             (setf (aref rows i) (triplet-row triplet)
                   (aref cols i) (triplet-col triplet)
                   (aref xs   i) (coerce (triplet-value triplet) 'double-float)))
    (sb-sys:with-pinned-objects (rows cols xs)
      (let ((sparse (nes-triplet-to-sparse
                     (sb-alien:sap-alien (sb-sys:vector-sap rows) (* int))
                     (sb-alien:sap-alien (sb-sys:vector-sap cols) (* int))
                     (sb-alien:sap-alien (sb-sys:vector-sap xs) (* double))
                     nnz nrow ncol common)))
        (when (null-alien sparse)
          (error "nes_triplet_to_sparse failed: ~A" (nes-last-error common)))
        sparse))))

(defun scale-sparse! (sparse scale                              ; :461-473
                      &aux (n (matlisp:nrows scale))
                        (common *cholmod-common*))
  (declare (type (alien (* cholmod-sparse)) sparse))
  (assert (= n (slot sparse 'ncol)))
  (with-store-saps ((s scale))
    (assert (/= (nes-scale s
                           2 ; scale columns
                           sparse common)
                0)))
  sparse)

(defun scale-sparse (sparse scale)                              ; :475-477
  (scale-sparse! (cholmod-copy-sparse sparse *cholmod-common*)
                 scale))

;;; ---- analyze / factorize / solve -----------------------------------------------------------------
(define-alien-routine ("nes_analyze" cholmod-analyze) (* cholmod-factor)               ; :261-263
  (a (* cholmod-sparse)) (common (* cholmod-common)))
(define-alien-routine ("nes_factorize" cholmod-factorize) int                          ; :265-268
  (a (* cholmod-sparse)) (l (* cholmod-factor)) (common (* cholmod-common)))
(define-alien-routine ("nes_solve" nes-solve) int                                      ; cholmod_solve :270-274
  (sys int) (l (* cholmod-factor)) (b (* double)) (x (* double)) (common (* cholmod-common)))
(define-alien-routine ("nes_solve2" nes-solve2) int                                    ; cholmod_solve2 :276-288
  (sys int) (l (* cholmod-factor)) (b (* double)) (x (* double)) (common (* cholmod-common)))
(define-alien-routine ("nes_free_factor" cholmod-free-factor) int                      ; :145-147
  (l (* (* cholmod-factor))) (common (* cholmod-common)))
(define-alien-routine ("nes_solve_dense" nes-solve-dense) int
  (b-matrix (* double)) (nrow size-t) (ncol size-t) (b (* double)) (x (* double))
  (common (* cholmod-common)))

;; Solve for (A A')x = b                                        ; :408-431
(defun solve-dense (A b &aux (common *cholmod-common*))
  (let ((x (matlisp:make-real-matrix-dim (matlisp:nrows A) 1)))
    (cholmod-set-status common 0)
    (let ((rc (with-store-saps ((pa A) (pb b) (px x))
                (nes-solve-dense pa (matlisp:nrows A) (matlisp:ncols A) pb px common))))
      (cond ((zerop rc) x)
            ((plusp rc) nil)                                    ; status 1: not positive definite -> NIL (:420-421)
            (t (error "nes_solve_dense: ~A" (nes-last-error common)))))))

(defstruct solve-sparse-state                                   ; :479-484
  factor
  rhs
  solution
  workspace-y
  workspace-e)

(defun free-sparse-state (state)                                ; :486-504
  ;; rhs / solution / workspaces live inside the nes-factor; only the factor is owned here
  (when (solve-sparse-state-factor state)
    (with-alien ((L (* cholmod-factor) :local
                    (shiftf (solve-sparse-state-factor state) nil)))
      (assert (/= 0 (cholmod-free-factor (addr L) *cholmod-common*)))))
  (setf (solve-sparse-state-rhs state) nil
        (solve-sparse-state-solution state) nil
        (solve-sparse-state-workspace-y state) nil
        (solve-sparse-state-workspace-e state) nil))

(defun solve-sparse-one-shot (As b &aux (common *cholmod-common*))   ; :506-522
  (with-alien ((As (* cholmod-sparse) :local As)
               (factor (* cholmod-factor) :local (cholmod-analyze As common)))
    (when (null-alien factor)
      (error "nes_analyze failed: ~A" (nes-last-error common)))
    (cholmod-set-status common 0)
    (cholmod-factorize As factor common)
    (when (/= (cholmod-get-status common) 0)
      (cholmod-free-factor (addr factor) common)              ; the reference leaks here (:513-514)
      (return-from solve-sparse-one-shot))
    (let ((x (matlisp:make-real-matrix-dim (matlisp:nrows b) 1)))
      (let ((rc (with-store-saps ((pb b) (px x))
                  (nes-solve 0 factor pb px common))))
        (cholmod-free-factor (addr factor) common)
        (unless (zerop rc)
          (error "nes_solve: ~A" (nes-last-error common))))
      x)))

(defun solve-sparse-recycle (As b state factorized             ; :524-560
                             &aux (common *cholmod-common*))
  (with-alien ((As (* cholmod-sparse) :local As)
               (factor (* cholmod-factor) :local
                       (or (solve-sparse-state-factor state)
                           (setf (solve-sparse-state-factor state)
                                 (cholmod-analyze As common)))))
    (unless factorized
      (cholmod-set-status common 0)
      (cholmod-factorize As factor common)
      (when (/= (cholmod-get-status common) 0)
        (return-from solve-sparse-recycle)))
    ;; solution in place of the right-hand side, like (dense-to-matlisp x matlisp-b) (:560);
    ;; the library uploads b before it writes x, so the two may alias
    (unless (plusp (with-store-saps ((pb b))
                     (nes-solve2 0 factor pb pb common)))
      (format t "status: ~A~%" (cholmod-get-status common)))
    b))

(defun solve-sparse (As b &optional state factorized)          ; :562-565
  (if state
      (solve-sparse-recycle As b state factorized)
      (solve-sparse-one-shot As b)))

(defun sparse-m* (sparse x &key transpose                      ; :567-614
                             y
                             (alpha 1d0)
                             (beta (if y 1d0 0d0))
                             output
                  &aux (nrow (slot sparse 'nrow))
                    (common *cholmod-common*))
  (declare (type (alien (* cholmod-sparse)) sparse))
  (let ((m nrow)
        (n (slot sparse 'ncol)))
    (when transpose (rotatef m n))
    (assert (= 1 (matlisp:ncols x)))
    (assert (= n (matlisp:nrows x)))
    (when y
      (assert (= 1 (matlisp:ncols y)))
      (assert (= m (matlisp:nrows y))))
    ;; The reference copies y into a fresh cholmod_dense (or zeros), multiplies into it and copies the
    ;; result into OUTPUT (or a fresh matrix): y itself is only overwritten when it is also OUTPUT.
    (let ((result (cond ((and output (eq output y)) y)
                        (output
                         (assert (= m (matlisp:nrows output)))
                         (assert (= 1 (matlisp:ncols output)))
                         (if y
                             (matlisp:copy! y output)
                             (matlisp:fill-matrix output 0d0))
                         output)
                        (y (matlisp:copy y))
                        (t (matlisp:make-real-matrix-dim m 1)))))
      (with-alien ((a (array double 2))
                   (b (array double 2)))
        (setf (deref a 0) (coerce alpha 'double-float)
              (deref a 1) 0d0
              (deref b 0) (coerce beta 'double-float)
              (deref b 1) 0d0)
        (unless (/= 0 (with-store-saps ((px x) (py result))
                        (nes-sdmult sparse (if transpose 1 0)
                                    (addr (deref a 0))
                                    (addr (deref b 0))
                                    px py common)))
          (flush)
          (error "nes_sdmult failed: ~A" (nes-last-error common))))
      result)))

;;; ---- optional: the whole reduction of solve-kkt-newton / the PDAS loop in one foreign call -----------
;;; (newton-solve.lisp:139-154, sparse-newton-solve.lisp:150-168, primal-dual-affine-scaling.lisp:319-396).
;;; Not needed for the reference's tests; these are the calls bench.py times.
(define-alien-routine ("nes_kkt_newton" nes-kkt-newton) int
  (a (* cholmod-sparse)) (l-factor (* cholmod-factor)) (filters int)
  (l (* double)) (u (* double)) (w (* double)) (z (* double))
  (e (* double)) (f (* double)) (g (* double)) (h (* double))
  (dw (* double)) (dx (* double)) (dy (* double)) (dz (* double))
  (common (* cholmod-common)))

(defun solve-kkt-newton/device (l u w z A e f g h &key (filters t) factor
                                &aux (common *cholmod-common*)
                                  (n (matlisp:nrows l)) (m (matlisp:nrows g)))
  "Same inputs and values as solve-kkt-newton; inputs are NOT destroyed.  NIL when the Cholesky fails;
signals DIVISION-BY-ZERO where the reference's scale-Z would (filter-Z, sparse-newton-solve.lisp:40-53)."
  (let ((dw (matlisp:make-real-matrix-dim n 1)) (dx (matlisp:make-real-matrix-dim n 1))
        (dy (matlisp:make-real-matrix-dim m 1)) (dz (matlisp:make-real-matrix-dim n 1)))
    (let ((rc (with-store-saps ((pl l) (pu u) (pw w) (pz z) (pe e) (pf f) (pg g) (ph h)
                                (pdw dw) (pdx dx) (pdy dy) (pdz dz))
                (nes-kkt-newton A (or factor (sb-alien:sap-alien (sb-sys:int-sap 0) (* cholmod-factor)))
                                (if filters 1 0)
                                pl pu pw pz pe pf pg ph pdw pdx pdy pdz common))))
      (case rc
        (0 (values dw dx dy dz))
        (1 nil)
        (2 (error 'division-by-zero :operation '/ :operands (list 'l 'z)))
        (t (error "nes_kkt_newton: ~A" (nes-last-error common)))))))

) ; #-nes-after-affine-scaling

;;; ---- the one override: load after affine-scaling.lisp ------------------------------------------------
#+nes-after-affine-scaling
(defun affine-A-copy (affine)                                   ; affine-scaling.lisp:29-35
  ;; The reference restores the scaled copy by memcpy'ing A's value array over it.  Here a copy shares
  ;; A's immutable device values and only carries a column scale: restoring = dropping the scale.
  (let ((A (affine-A affine))
        (copy (affine-%a-copy affine)))
    (declare (type (alien (* cholmod-sparse)) A copy)
             (ignorable A))
    (nes-unscale copy *cholmod-common*)
    copy))
